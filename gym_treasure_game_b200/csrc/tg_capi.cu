// C ABI of libtreasure_b200.so (include/treasure_b200.h): level compiler, batch
// lifetime, and the stream-ordered entry points.  Host code only; kernels live in
// tg_step.cu / tg_render.cu.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <algorithm>
#include <climits>
#include <cstring>
#include <new>
#include <vector>

#include <omp.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "tg_launch.h"
#include "tg_host_patch.h"

using namespace tg;

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(TG_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------
// objects behind the opaque handles
// ---------------------------------------------------------------------------
struct tg_level {
    LevelBlob blob;
    tg_level_info info;
    std::vector<uint8_t> background;   // frame_h * frame_w * 3
    std::vector<uint32_t> sprites;     // TG_NUM_SPRITES * 48 * 48
    std::vector<uint32_t> closure;     // LevelBlob::closure, built once per level
};

struct tg_env {
    int device = 0;
    int n_levels = 0;
    int ni = 2;                        // kernel specialisation: 2 or 4 item slots
    BatchView B{};
    RenderView R{};
    bool has_render = false;
    std::vector<void *> allocs;        // everything cudaMalloc'ed for this env
    // device-side staging for tg_step_host (owned through `allocs`)
    int32_t *s_actions = nullptr; float *s_obs = nullptr; float *s_reward = nullptr;
    uint8_t *s_done = nullptr; uint8_t *s_ran = nullptr;
    // Observation bookkeeping: the step kernel only writes the observation rows of envs whose state changed (an option
    // that cannot run leaves its env untouched, opt:22-23).  `obs_sync` is the device buffer known to hold every env's
    // current row (NULL: none); it is trusted when it is the caller's bound buffer (tg_bind_obs) or the library's own
    // staging buffer.  Any other `obs` argument gets every row written (tg_obs_kernel after the step kernel).
    const float *obs_bound = nullptr, *obs_sync = nullptr;
    uint4 *tr_core = nullptr; uint2 *tr_items = nullptr; double *tr_angles = nullptr; uint8_t *tr_lid = nullptr;   // tg_step_frames snapshots
    int64_t tr_slots = 0;
    int step_tile = 0;                 // envs per step-kernel CTA; 0 = chosen by the launcher (TG_STEP_TILE / tg_debug_set_step_tile)
    int64_t launches = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;   // copied by the *_host entry points
    // tg_step_host pipeline: step kernel of chunk c+1 overlaps the device->host copies of chunk c
    cudaStream_t side = nullptr;
    cudaEvent_t ev_chunk[8] = {}, ev_join = nullptr;
    // tg_step_host_sparse: device / pinned-host record buffers, per-chunk counters, and what the caller's host arrays
    // are known to hold (they are only patched while `sp_primed` and the pointers have not changed)
    uint32_t *sp_drecs = nullptr;                             // per chunk: 16-byte header (record count) + records
    uint32_t *sp_hrecs = nullptr;                             // the same regions in pinned host memory (cudaHostAlloc)
    cudaEvent_t ev_rec[8] = {}, ev_h2d[8] = {};
    size_t sp_off[8] = {}, sp_table_off = 0;                  // region offsets (words) of the chunks in flight
    int sp_grid[8] = {}, sp_tile[8] = {};                     // CTAs and envs per CTA of each chunk's launch
    int64_t sp_copied[8] = {}, sp_per = 0;                    // between tg_step_host_sparse_begin and _end: records requested per chunk, chunk size
    int sp_pending = 0, sp_used = 0;                          // 1 = a sparse step is in flight, 2 = begin ran the dense path; chunks in flight
    int64_t sp_guess[8] = {};                                 // records to copy with the header (previous count + margin), < 0 = all
    int sp_words = 0;
    bool sp_primed = false;
    double sp_t_enqueue = 0, sp_t_wait = 0, sp_t_patch = 0;   // seconds spent by the sparse host step: enqueueing, waiting for records, patching
    const void *sp_obs = nullptr, *sp_reward = nullptr, *sp_done = nullptr, *sp_ran = nullptr;
    // mixed-layout batches: per layout the ascending list of its env indices (host copy for the range search, device copy
    // for the renderer, which draws a mixed batch layout by layout)
    std::vector<int32_t> lev_envs[TG_MAX_LEVELS];
    int32_t *d_lev_envs[TG_MAX_LEVELS] = {};
};

extern "C" const char *tg_last_error(void) { return g_err; }
extern "C" int tg_abi_version(void) { return TG_ABI_VERSION; }
extern "C" int tg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---------------------------------------------------------------------------
// level compiler (host)
// ---------------------------------------------------------------------------
// LevelBlob::closure: set_val (objs:145-149, :175-178, :231-235) followed by process_trigger (objs:76-94) for every
// (object o, value v, door / handle / bolt bits), for an env whose handles carry no sticky previously_triggered flag.  The
// walk below is the device's trigger_dfs (tg_device.cuh) step for step: same lists, same order, same blocking set; instead
// of drawing it records the handles whose angle is redrawn.
static void build_closure(const LevelBlob &b, std::vector<uint32_t> &tab) {
    const int nobj = b.n_objs;
    const uint32_t NB = 1u << CLOSURE_BITS;
    tab.assign((size_t)nobj * 2 * NB, 7u << CLOSURE_BITS);                // not tabulated
    auto bit_of = [&](int o) {
        const int k = b.obj_kind[o], i = b.obj_idx[o];
        return (k == TG_DOOR) ? i : (k == TG_HANDLE) ? TG_MAX_DOORS + i : TG_MAX_DOORS + TG_MAX_HANDLES + i;
    };
    uint32_t valid = 0;
    for (int o = 0; o < nobj; o++) if (b.obj_kind[o] == TG_DOOR || b.obj_kind[o] == TG_HANDLE || b.obj_kind[o] == TG_BOLT) valid |= 1u << bit_of(o);
    for (int o0 = 0; o0 < nobj; o0++) {
        if (!(b.obj_kind[o0] == TG_DOOR || b.obj_kind[o0] == TG_HANDLE || b.obj_kind[o0] == TG_BOLT)) continue;
        for (int v0 = 0; v0 < 2; v0++)
            for (uint32_t f0 = 0; f0 < NB; f0++) {
                if (f0 & ~valid) continue;
                uint32_t f = f0, events = 0;
                int ne = 0;
                bool overflow = false;
                auto apply = [&](int o, bool v) {
                    const int bit = bit_of(o);
                    if ((bool)((f >> bit) & 1u) == v) return false;
                    f ^= 1u << bit;
                    if (b.obj_kind[o] == TG_HANDLE) {
                        if (ne < 5) events |= ((uint32_t)b.obj_idx[o] | (v ? 4u : 0u)) << (3 * ne);
                        else overflow = true;
                        ne++;
                    }
                    return true;
                };
                if (apply(o0, v0 != 0)) {
                    int st_obj[TG_MAX_OBJECTS + 1], st_t[TG_MAX_OBJECTS + 1], sp = 0;
                    uint32_t pt = 1u << o0;
                    st_obj[0] = o0; st_t[0] = b.trig_begin[2 * o0 + v0]; sp = 1;
                    while (sp > 0) {
                        const int l = sp - 1, o = st_obj[l];
                        int t = st_t[l];
                        const int tend = b.trig_begin[2 * o + (int)((f >> bit_of(o)) & 1u) + 1];
                        bool pushed = false;
                        while (t < tend) {
                            const int ent = b.trig_list[t++];
                            const int dst = ent & 127; const bool dv = (ent >> 7) != 0;
                            if (pt & (1u << dst)) continue;
                            if (apply(dst, dv)) {
                                st_t[l] = t;
                                pt |= 1u << dst;
                                st_obj[sp] = dst; st_t[sp] = b.trig_begin[2 * dst + (dv ? 1 : 0)];
                                sp++;
                                pushed = true;
                                break;
                            }
                        }
                        if (!pushed) { pt &= ~(1u << o); sp--; }
                    }
                }
                if (!overflow) tab[(((size_t)o0 * 2 + v0) << CLOSURE_BITS) | f0] = f | ((uint32_t)ne << CLOSURE_BITS) | (events << 16);
            }
    }
}

extern "C" int tg_level_create(const uint8_t *tiles, int32_t cw, int32_t ch, const tg_object *objs, int32_t n_objs,
                               const tg_trigger *trigs, int32_t n_trigs, tg_level **out) {
    if (!tiles || !out || (n_objs > 0 && !objs) || (n_trigs > 0 && !trigs)) return fail(TG_ERR_ARG, "null argument");
    if (cw < 1 || ch < 1 || cw > TG_MAX_GRID || ch > TG_MAX_GRID) return fail(TG_ERR_ARG, "grid %dx%d outside 1..%d", cw, ch, TG_MAX_GRID);
    if (n_objs < 0 || n_objs > TG_MAX_OBJECTS) return fail(TG_ERR_ARG, "%d objects (max %d)", n_objs, TG_MAX_OBJECTS);
    if (n_trigs < 0 || n_trigs > TG_MAX_TRIGGERS) return fail(TG_ERR_ARG, "%d triggers (max %d)", n_trigs, TG_MAX_TRIGGERS);
    tg_level *lv = new (std::nothrow) tg_level();
    if (!lv) return fail(TG_ERR_NOMEM, "out of host memory");
    LevelBlob &b = lv->blob;
    memset(&b, 0, sizeof b);
    memset(b.tiles, T_WALL, sizeof b.tiles);                 // out of bounds reads as WALL (impl:220-223)
    b.cw = (int16_t)cw; b.ch = (int16_t)ch;
    b.start_px = b.start_py = -1;
    for (int y = 0; y < ch; y++)
        for (int x = 0; x < cw; x++) {
            const uint8_t c = tiles[y * cw + x];
            int t;
            if (c == ' ') t = T_OPEN; else if (c == '/') t = T_WALL; else if (c == 'L') t = T_LADDER;
            else { delete lv; return fail(TG_ERR_ARG, "tile (%d,%d) has unsupported character 0x%02x", x, y, c); }
            b.tiles[(y + PAD) * TSTRIDE + x + PAD] = (uint8_t)t;
            if (t != T_WALL && b.start_px < 0) { b.start_px = (int16_t)(x * S + S / 2); b.start_py = (int16_t)(y * S); lv->info.start_cx = x; lv->info.start_cy = y; }
        }
    if (b.start_px < 0) { lv->info.start_cx = lv->info.start_cy = -1; }
    uint32_t flags = 1u << F_FACING;
    int obs = 2;
    memset(b.obs_prog, OP_ZERO << 4, sizeof b.obs_prog);
    b.obs_prog[0] = OP_PX << 4; b.obs_prog[1] = OP_PY << 4;
    int obj_of[5][TG_MAX_OBJECTS];
    int cnt[5] = {0, 0, 0, 0, 0};
    int n_items = 0;
    for (int o = 0; o < n_objs; o++) {
        const tg_object &ob = objs[o];
        if (ob.kind < 0 || ob.kind > 4) { delete lv; return fail(TG_ERR_ARG, "object %d: unknown kind %d", o, ob.kind); }
        if (ob.cx < 0 || ob.cx >= cw || ob.cy < 0 || ob.cy >= ch) { delete lv; return fail(TG_ERR_ARG, "object %d outside the grid", o); }
        uint8_t &code = b.tiles[(ob.cy + PAD) * TSTRIDE + ob.cx + PAD];
        b.obj_kind[o] = (uint8_t)ob.kind;
        b.obj_obs[o] = 255;
        obj_of[ob.kind][cnt[ob.kind]] = o;
        int idx;
        switch (ob.kind) {
        case TG_DOOR:
            idx = b.n_doors;
            if (idx >= TG_MAX_DOORS) { delete lv; return fail(TG_ERR_ARG, "more than %d doors", TG_MAX_DOORS); }
            if (code & TC_HAS_DOOR) { delete lv; return fail(TG_ERR_ARG, "two doors share cell (%d,%d)", ob.cx, ob.cy); }
            code = (uint8_t)((code & TC_STATIC_OBJ) | TC_HAS_DOOR | (idx << 4));
            b.door_cx[idx] = (int8_t)ob.cx; b.door_cy[idx] = (int8_t)ob.cy; b.door_obj[idx] = (uint8_t)o;
            if (ob.flag) flags |= 1u << (F_DOORS + idx);
            b.n_doors++;
            break;
        case TG_HANDLE:
            idx = b.n_handles;
            if (idx >= TG_MAX_HANDLES) { delete lv; return fail(TG_ERR_ARG, "more than %d handles", TG_MAX_HANDLES); }
            code |= TC_STATIC_OBJ;
            b.handle_cx[idx] = (int8_t)ob.cx; b.handle_cy[idx] = (int8_t)ob.cy; b.handle_obj[idx] = (uint8_t)o;
            if (ob.flag) flags |= 1u << (F_HANDLES + idx);
            b.obj_obs[o] = (uint8_t)obs; b.obs_prog[obs] = (uint8_t)(OP_ANGLE << 4 | idx); obs += 1;
            b.n_handles++;
            break;
        case TG_BOLT:
            idx = b.n_bolts;
            if (idx >= TG_MAX_BOLTS) { delete lv; return fail(TG_ERR_ARG, "more than %d bolts", TG_MAX_BOLTS); }
            code |= TC_STATIC_OBJ;
            b.bolt_cx[idx] = (int8_t)ob.cx; b.bolt_cy[idx] = (int8_t)ob.cy; b.bolt_obj[idx] = (uint8_t)o;
            if (ob.flag) flags |= 1u << (F_BOLTS + idx);
            b.obj_obs[o] = (uint8_t)obs; b.obs_prog[obs] = (uint8_t)(OP_BOLT << 4 | idx); obs += 1;
            b.n_bolts++;
            break;
        default:   // key, gold
            idx = n_items;
            if (idx >= TG_MAX_ITEMS) { delete lv; return fail(TG_ERR_ARG, "more than %d keys+gold", TG_MAX_ITEMS); }
            b.item_cx[idx] = (int8_t)ob.cx; b.item_cy[idx] = (int8_t)ob.cy; b.item_obj[idx] = (uint8_t)o;
            if (ob.kind == TG_KEY) b.key_mask |= (uint8_t)(1u << idx); else b.gold_mask |= 1u << idx;
            b.obj_obs[o] = (uint8_t)obs;
            b.obs_prog[obs] = (uint8_t)(OP_IX << 4 | idx); b.obs_prog[obs + 1] = (uint8_t)(OP_IY << 4 | idx); obs += 2;
            n_items++;
            break;
        }
        b.obj_idx[o] = (uint8_t)idx;
        cnt[ob.kind]++;
    }
    b.n_items = (uint8_t)n_items; b.n_objs = (uint8_t)n_objs; b.obs_dim = (uint8_t)obs;
    {   // obs_prog[31] = 1: slot order of the shipped level (write_obs has straight-line code for it)
        static const uint8_t shipped[9] = {OP_PX << 4, OP_PY << 4, OP_ANGLE << 4 | 0, OP_ANGLE << 4 | 1, OP_IX << 4 | 0,
                                           OP_IY << 4 | 0, OP_BOLT << 4 | 0, OP_IX << 4 | 1, OP_IY << 4 | 1};
        b.obs_prog[31] = (obs == 9 && memcmp(b.obs_prog, shipped, 9) == 0) ? 1 : 0;
    }
    b.init_flags = flags;
    // bag slots are the cells (cw-1-k, ch-1) (impl:353); they must be unreachable, i.e. plain WALL,
    // so that an item in the bag can never be picked up a second time
    for (int k = 0; k < n_items; k++) {
        const int bx = cw - 1 - k, by = ch - 1;
        if (bx < 0 || b.tiles[(by + PAD) * TSTRIDE + bx + PAD] != T_WALL) {
            delete lv;
            return fail(TG_ERR_ARG, "bag cell (%d,%d) must be a wall cell without objects", bx, by);
        }
    }
    for (int t = 0; t < n_trigs; t++) {
        const tg_trigger &tr = trigs[t];
        const bool ok1 = tr.src_kind == TG_DOOR || tr.src_kind == TG_HANDLE || tr.src_kind == TG_BOLT;
        const bool ok2 = tr.dst_kind == TG_DOOR || tr.dst_kind == TG_HANDLE || tr.dst_kind == TG_BOLT;
        if (!ok1 || !ok2 || tr.src_index < 0 || tr.src_index >= cnt[tr.src_kind] || tr.dst_index < 0 || tr.dst_index >= cnt[tr.dst_kind]) {
            delete lv;
            return fail(TG_ERR_ARG, "trigger %d references a missing object", t);
        }
        b.trig_src[t] = (uint8_t)(obj_of[tr.src_kind][tr.src_index] | (tr.src_value ? 128 : 0));
        b.trig_dst[t] = (uint8_t)(obj_of[tr.dst_kind][tr.dst_index] | (tr.dst_value ? 128 : 0));
    }
    b.n_trigs = (uint8_t)n_trigs;
    {   // grouped by (source object, value), file order inside a group
        int pos = 0;
        for (int key = 0; key < 2 * TG_MAX_OBJECTS; key++) {
            b.trig_begin[key] = (uint8_t)pos;
            const uint8_t src = (uint8_t)((key >> 1) | ((key & 1) ? 128 : 0));
            for (int t = 0; t < n_trigs; t++) if (b.trig_src[t] == src) b.trig_list[pos++] = b.trig_dst[t];
        }
        for (int key = 2 * TG_MAX_OBJECTS; key < 2 * TG_MAX_OBJECTS + 8; key++) b.trig_begin[key] = (uint8_t)pos;
    }
    // row bit masks (door cells read as OPEN here; closed doors are OR-ed in from door_lut at run time)
    for (int r = 0; r < TSTRIDE; r++)
        for (int c = 0; c < TSTRIDE; c++) {
            const uint8_t code = b.tiles[r * TSTRIDE + c];
            if (code & TC_STATIC_OBJ) b.row_static_obj[r] |= 1u << c;
            if (code & TC_HAS_DOOR) continue;
            const int t = code & 3;
            if (t != T_OPEN) { b.row_nonopen[r] |= 1u << c; b.col_nonopen[c] |= 1u << r; }
            if (t == T_WALL) b.row_solid[r] |= 1u << c;
            if (t == T_LADDER) { b.row_ladder[r] |= 1u << c; b.col_ladder[c] |= 1u << r; }
        }
    {
        int n_rows = 0;
        for (int d = 0; d < b.n_doors; d++) {
            const int r = b.door_cy[d] + PAD;
            if (!b.row_lut[r]) b.row_lut[r] = (uint8_t)(++n_rows);          // at most TG_MAX_DOORS distinct rows
        }
        for (int d = 0; d < b.n_doors; d++) {
            const int li = b.row_lut[b.door_cy[d] + PAD] - 1;
            for (int closed = 0; closed < 64; closed++)
                if ((closed >> d) & 1) b.door_lut[li][closed] |= 1u << (b.door_cx[d] + PAD);
        }
    }
    b.inv_w = 1.0f / (float)(cw * S); b.inv_h = 1.0f / (float)(ch * S);
    tg_level_info &I = lv->info;
    I.cw = cw; I.ch = ch; I.n_doors = b.n_doors; I.n_handles = b.n_handles; I.n_bolts = b.n_bolts; I.n_items = n_items;
    I.n_objects = n_objs; I.n_triggers = n_trigs; I.obs_dim = obs; I.frame_w = cw * S; I.frame_h = ch * S; I.has_sprites = 0;
    build_closure(b, lv->closure);
    *out = lv;
    return TG_OK;
}

extern "C" int tg_level_get_info(const tg_level *lv, tg_level_info *out) {
    if (!lv || !out) return fail(TG_ERR_ARG, "null argument");
    *out = lv->info;
    return TG_OK;
}

extern "C" int tg_level_set_sprites(tg_level *lv, const uint8_t *sprites_rgba, const uint8_t *background_rgb) {
    if (!lv || !sprites_rgba || !background_rgb) return fail(TG_ERR_ARG, "null argument");
    const size_t nb = (size_t)lv->info.frame_w * lv->info.frame_h * 3;
    lv->background.assign(background_rgb, background_rgb + nb);
    lv->sprites.resize((size_t)TG_NUM_SPRITES * S * S);
    memcpy(lv->sprites.data(), sprites_rgba, lv->sprites.size() * 4);   // bytes R,G,B,A == little-endian R|G<<8|B<<16|A<<24
    lv->info.has_sprites = 1;
    return TG_OK;
}

extern "C" void tg_level_destroy(tg_level *lv) { delete lv; }

// ---------------------------------------------------------------------------
// batch lifetime
// ---------------------------------------------------------------------------
template <typename T>
static cudaError_t dev_alloc(tg_env *e, T **p, size_t count) {
    void *q = nullptr;
    cudaError_t r = cudaMalloc(&q, count * sizeof(T) ? count * sizeof(T) : 16);
    if (r == cudaSuccess) { e->allocs.push_back(q); *p = static_cast<T *>(q); }
    return r;
}

static void free_env(tg_env *e) {
    if (!e) return;
    if (e->side) cudaStreamDestroy(e->side);
    for (cudaEvent_t ev : e->ev_chunk) if (ev) cudaEventDestroy(ev);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    for (cudaEvent_t ev : e->ev_rec) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->ev_h2d) if (ev) cudaEventDestroy(ev);
    if (e->sp_hrecs) cudaFreeHost(e->sp_hrecs);
    for (void *p : {(void *)e->tr_core, (void *)e->tr_items, (void *)e->tr_angles, (void *)e->tr_lid}) if (p) cudaFree(p);
    for (void *p : e->allocs) cudaFree(p);
    delete e;
}

extern "C" int tg_create(const tg_level *const *levels, int32_t n_levels, const uint8_t *level_ids, int64_t num_envs,
                         int64_t first_env_id, int32_t device, uint64_t seed, int32_t max_episode_steps,
                         int32_t auto_reset, tg_env **out) {
    if (!levels || !out) return fail(TG_ERR_ARG, "null argument");
    if (n_levels < 1 || n_levels > TG_MAX_LEVELS) return fail(TG_ERR_ARG, "n_levels %d outside 1..%d", n_levels, TG_MAX_LEVELS);
    if (num_envs < 1) return fail(TG_ERR_ARG, "num_envs must be positive");
    if (max_episode_steps < 0) return fail(TG_ERR_ARG, "max_episode_steps must be >= 0");
    for (int l = 0; l < n_levels; l++) if (!levels[l]) return fail(TG_ERR_ARG, "level %d is null", l);
    if (level_ids)
        for (int64_t i = 0; i < num_envs; i++)
            if (level_ids[i] >= n_levels) return fail(TG_ERR_ARG, "level_ids[%lld] = %d out of range", (long long)i, level_ids[i]);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(TG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(TG_ERR_ARG, "device %d outside 0..%d", device, ndev - 1);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(TG_ERR_CUDA, "cannot select device %d", device);

    tg_env *e = new (std::nothrow) tg_env();
    if (!e) return fail(TG_ERR_NOMEM, "out of host memory");
    e->device = device; e->n_levels = n_levels;
    { const char *v = getenv("TG_STEP_TILE"); e->step_tile = v ? atoi(v) : 0; }
    BatchView &B = e->B;
    B.n = num_envs; B.first_env_id = first_env_id; B.n_levels = n_levels;
    B.r_begin = 0; B.r_count = num_envs;
    B.max_steps = max_episode_steps; B.auto_reset = auto_reset ? 1 : 0;
    B.seed_lo = (uint32_t)seed; B.seed_hi = (uint32_t)(seed >> 32);
    int max_items = 0, obs_dim = 0;
    bool all_sprites = true, same_size = true;
    for (int l = 0; l < n_levels; l++) {
        if (levels[l]->info.n_items > max_items) max_items = levels[l]->info.n_items;
        if (levels[l]->info.obs_dim > obs_dim) obs_dim = levels[l]->info.obs_dim;
        all_sprites = all_sprites && levels[l]->info.has_sprites;
        same_size = same_size && levels[l]->info.cw == levels[0]->info.cw && levels[l]->info.ch == levels[0]->info.ch;
    }
    e->ni = (max_items <= 2) ? 2 : 4;
    B.obs_dim = obs_dim;

#define CUE(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { free_env(e); return fail(TG_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)

    LevelBlob *d_levels = nullptr;
    CUE(dev_alloc(e, &d_levels, (size_t)n_levels));
    const bool no_closure = getenv("TG_NO_CLOSURE") != nullptr;            // debug / tests: every interact walks the trigger graph
    for (int l = 0; l < n_levels; l++) {
        LevelBlob blob = levels[l]->blob;
        if (!levels[l]->closure.empty() && !no_closure) {               // the trigger closure lives in global memory, the blob points to it
            uint32_t *d_cl = nullptr;
            CUE(dev_alloc(e, &d_cl, levels[l]->closure.size()));
            CUE(cudaMemcpy(d_cl, levels[l]->closure.data(), levels[l]->closure.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            blob.closure = d_cl;
        }
        CUE(cudaMemcpy(d_levels + l, &blob, sizeof(LevelBlob), cudaMemcpyHostToDevice));
    }
    B.levels = d_levels;
    CUE(dev_alloc(e, &B.core, (size_t)num_envs));
    CUE(dev_alloc(e, &B.acct, (size_t)num_envs));
    CUE(dev_alloc(e, &B.plan, (size_t)num_envs + 4));       // + 4: the step kernel's 16-byte loads never leave the allocation
    CUE(dev_alloc(e, &B.ep_start, (size_t)num_envs + 4));
    CUE(cudaMemset(B.plan, 0, sizeof(uint64_t) * ((size_t)num_envs + 4)));
    CUE(cudaMemset(B.ep_start, 0, sizeof(uint32_t) * ((size_t)num_envs + 4)));
    CUE(dev_alloc(e, &B.step_counter, (size_t)4));
    CUE(cudaMemset(B.step_counter, 0, 4 * sizeof(uint32_t)));
    B.advance = 1;
    CUE(cudaMemset(B.core, 0, sizeof(uint4) * (size_t)num_envs));
    CUE(cudaMemset(B.acct, 0, sizeof(uint4) * (size_t)num_envs));
    if (e->ni > 2) { CUE(dev_alloc(e, &B.items23, (size_t)num_envs)); CUE(cudaMemset(B.items23, 0, sizeof(uint2) * (size_t)num_envs)); }
    CUE(dev_alloc(e, &B.angles, (size_t)num_envs * TG_MAX_HANDLES));
    CUE(cudaMemset(B.angles, 0, sizeof(double) * (size_t)num_envs * TG_MAX_HANDLES));
    if (level_ids && n_levels > 1) {
        uint8_t *d_ids = nullptr;
        CUE(dev_alloc(e, &d_ids, (size_t)num_envs));
        CUE(cudaMemcpy(d_ids, level_ids, (size_t)num_envs, cudaMemcpyHostToDevice));
        B.level_id = d_ids;
        if (num_envs <= INT32_MAX) {
            for (int64_t i = 0; i < num_envs; i++) e->lev_envs[level_ids[i]].push_back((int32_t)i);
            for (int l = 0; l < n_levels; l++) {
                if (e->lev_envs[l].empty()) continue;
                CUE(dev_alloc(e, &e->d_lev_envs[l], e->lev_envs[l].size()));
                CUE(cudaMemcpy(e->d_lev_envs[l], e->lev_envs[l].data(), e->lev_envs[l].size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            }
        }
    }
    {   // quotient tables of the observation (impl:368-378): float32 of the reference's float64 x / width, y / height
        std::vector<float> lut((size_t)n_levels * 2 * OBS_LUT_N);
        for (int l = 0; l < n_levels; l++) {
            const double W = (double)levels[l]->info.frame_w, H = (double)levels[l]->info.frame_h;
            for (int v = 0; v < OBS_LUT_N; v++) {
                lut[((size_t)l * 2 + 0) * OBS_LUT_N + v] = (float)((double)(v - S) / W);
                lut[((size_t)l * 2 + 1) * OBS_LUT_N + v] = (float)((double)(v - S) / H);
            }
        }
        float *d_lut = nullptr;
        CUE(dev_alloc(e, &d_lut, lut.size()));
        CUE(cudaMemcpy(d_lut, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice));
        B.obs_lut = d_lut;
    }
    CUE(dev_alloc(e, &B.stats, (size_t)8));
    CUE(cudaMemset(B.stats, 0, 64));
    if (all_sprites && same_size) {
        RenderView &R = e->R;
        R.frame_w = levels[0]->info.frame_w; R.frame_h = levels[0]->info.frame_h;
        R.cw = levels[0]->info.cw; R.ch = levels[0]->info.ch;
        for (int l = 0; l < n_levels; l++) {
            uint8_t *bg = nullptr; uint32_t *sp = nullptr;
            CUE(dev_alloc(e, &bg, levels[l]->background.size()));
            CUE(cudaMemcpy(bg, levels[l]->background.data(), levels[l]->background.size(), cudaMemcpyHostToDevice));
            CUE(dev_alloc(e, &sp, levels[l]->sprites.size()));
            CUE(cudaMemcpy(sp, levels[l]->sprites.data(), levels[l]->sprites.size() * 4, cudaMemcpyHostToDevice));
            R.assets[l].background = bg; R.assets[l].sprites = sp;
        }
        CUE(dev_alloc(e, &R.job_counter, (size_t)1));
        e->has_render = true;
    }
    // constructor draws + initial state (impl:31-53)
    CUE(launch_reset(B, e->ni, nullptr, nullptr, 0));
    e->launches++;
    CUE(cudaDeviceSynchronize());
#undef CUE
    *out = e;
    return TG_OK;
}

extern "C" void tg_destroy(tg_env *env) {
    if (!env) return;
    DeviceGuard guard(env->device);
    cudaDeviceSynchronize();
    free_env(env);
}

extern "C" int64_t tg_num_envs(const tg_env *env) { return env ? env->B.n : 0; }
extern "C" int32_t tg_obs_dim(const tg_env *env) { return env ? env->B.obs_dim : 0; }
extern "C" int64_t tg_launch_count(const tg_env *env) { return env ? env->launches : 0; }
extern "C" void tg_host_traffic(const tg_env *env, int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (h2d_bytes) *h2d_bytes = env ? env->h2d_bytes : 0;
    if (d2h_bytes) *d2h_bytes = env ? env->d2h_bytes : 0;
}

// ---------------------------------------------------------------------------
// stream-ordered entry points
// ---------------------------------------------------------------------------
extern "C" int tg_reset(tg_env *env, const uint8_t *mask, float *obs, void *stream) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_reset between tg_step_host_sparse_begin and _end");
    DeviceGuard guard(env->device);
    CU(launch_reset(env->B, env->ni, mask, obs, (cudaStream_t)stream));
    env->launches++;
    env->sp_primed = false;
    env->obs_sync = obs;                 // the reset kernel writes every env's row (NULL: no buffer is current)
    return TG_OK;
}

static bool obs_is_current(const tg_env *env, const float *obs) {
    return obs && obs == env->obs_sync && (obs == env->obs_bound || obs == env->s_obs);
}

extern "C" int tg_step(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done, uint8_t *ran,
                       uint16_t *avail, void *stream) {
    if (!env || !actions) return fail(TG_ERR_ARG, "null argument");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_step between tg_step_host_sparse_begin and _end");
    DeviceGuard guard(env->device);
    CU(launch_step(env->B, env->ni, env->step_tile, actions, obs, reward, done, ran, avail, (cudaStream_t)stream));
    env->launches++;
    if (obs && !obs_is_current(env, obs)) {        // not the bound buffer with last step's rows in it: every row
        CU(launch_obs(env->B, env->ni, obs, (cudaStream_t)stream));
        env->launches++;
    }
    env->obs_sync = obs;
    env->sp_primed = false;
    return TG_OK;
}

extern "C" int tg_bind_flags(tg_env *env, uint8_t *done01, uint8_t *terminated01, uint8_t *truncated01) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    const int given = (done01 ? 1 : 0) + (terminated01 ? 1 : 0) + (truncated01 ? 1 : 0);
    if (given != 0 && given != 3) return fail(TG_ERR_ARG, "bind all three flag arrays or none");
    if (given && ((reinterpret_cast<uintptr_t>(done01) | reinterpret_cast<uintptr_t>(terminated01) | reinterpret_cast<uintptr_t>(truncated01)) & 3u))
        return fail(TG_ERR_ARG, "flag arrays must be 4-byte aligned");
    env->B.flag_done = done01; env->B.flag_term = terminated01; env->B.flag_trunc = truncated01;
    return TG_OK;
}

extern "C" int tg_bind_obs(tg_env *env, float *obs) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    env->obs_bound = obs;
    return TG_OK;
}

static int ensure_staging(tg_env *env) {
    if (env->s_ran) return TG_OK;
    const size_t n = (size_t)env->B.n;
    CU(dev_alloc(env, &env->s_actions, n));
    CU(dev_alloc(env, &env->s_obs, n * env->B.obs_dim));
    CU(dev_alloc(env, &env->s_reward, n));
    CU(dev_alloc(env, &env->s_done, n));
    CU(dev_alloc(env, &env->s_ran, n));
    return TG_OK;
}

extern "C" int tg_step_host(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                            uint8_t *ran, void *stream) {
    if (!env || !actions) return fail(TG_ERR_ARG, "null argument");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_step_host between tg_step_host_sparse_begin and _end");
    DeviceGuard guard(env->device);
    int rc = ensure_staging(env);
    if (rc != TG_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = env->B.n;
    const int od = env->B.obs_dim;
    // Large batches are split into chunks: while the copy engine drains chunk c's results over PCIe,
    // the SMs already run chunk c+1.  Env ranges are independent, so the result is the same as one launch.
    int chunks = 1;
    if (n >= (int64_t)4 * 131072) chunks = 4; else if (n >= (int64_t)2 * 131072) chunks = 2;
    if (chunks > 1 && !env->side) {
        CU(cudaStreamCreateWithFlags(&env->side, cudaStreamNonBlocking));
        for (int c = 0; c < 8; c++) CU(cudaEventCreateWithFlags(&env->ev_chunk[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&env->ev_join, cudaEventDisableTiming));
    }
    CU(cudaMemcpyAsync(env->s_actions, actions, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    env->h2d_bytes += n * (int64_t)sizeof(int32_t);
    env->d2h_bytes += n * ((obs ? (int64_t)od * 4 : 0) + (reward ? 4 : 0) + (done ? 1 : 0) + (ran ? 1 : 0));
    const int64_t per = ((n + chunks - 1) / chunks + 2047) / 2048 * 2048;      // tile-aligned chunk size
    const int used_chunks = (int)((n + per - 1) / per);
    const bool obs_full = obs && !obs_is_current(env, env->s_obs);             // the staging buffer lacks the rows of idle envs
    for (int c = 0; c < used_chunks; c++) {
        const int64_t lo = (int64_t)c * per, cnt = (lo + per <= n) ? per : n - lo;
        BatchView V = env->B;
        V.r_begin = lo; V.r_count = cnt; V.advance = (c == used_chunks - 1) ? 1 : 0;
        CU(launch_step(V, env->ni, env->step_tile, env->s_actions, obs ? env->s_obs : nullptr, reward ? env->s_reward : nullptr,
                       done ? env->s_done : nullptr, ran ? env->s_ran : nullptr, nullptr, s));
        env->launches++;
        if (obs_full) { CU(launch_obs(V, env->ni, env->s_obs, s)); env->launches++; }
        cudaStream_t cs = s;
        if (chunks > 1) {
            CU(cudaEventRecord(env->ev_chunk[c], s));
            CU(cudaStreamWaitEvent(env->side, env->ev_chunk[c], 0));
            cs = env->side;
        }
        if (obs) CU(cudaMemcpyAsync(obs + lo * od, env->s_obs + lo * od, (size_t)cnt * od * sizeof(float), cudaMemcpyDeviceToHost, cs));
        if (reward) CU(cudaMemcpyAsync(reward + lo, env->s_reward + lo, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, cs));
        if (done) CU(cudaMemcpyAsync(done + lo, env->s_done + lo, (size_t)cnt, cudaMemcpyDeviceToHost, cs));
        if (ran) CU(cudaMemcpyAsync(ran + lo, env->s_ran + lo, (size_t)cnt, cudaMemcpyDeviceToHost, cs));
    }
    if (chunks > 1) {
        CU(cudaEventRecord(env->ev_join, env->side));
        CU(cudaStreamWaitEvent(s, env->ev_join, 0));
    }
    CU(cudaStreamSynchronize(s));
    env->obs_sync = obs ? env->s_obs : nullptr;
    // the caller's arrays now hold every env's outputs: tg_step_host_sparse may patch them from here on
    env->sp_primed = obs && reward && done && ran;
    env->sp_obs = obs; env->sp_reward = reward; env->sp_done = done; env->sp_ran = ran;
    return TG_OK;
}

// Record region of chunk c, on the device and mirrored in pinned host memory: a 16-byte header whose first word is the
// record count (the kernel's atomic), then the records.  One copy brings the header and as many records as the previous
// step produced plus a margin; the count then says whether a second copy is needed (it rarely is: the share of envs
// that run is stable from step to step).  This removed a count round trip per chunk (0.17 of 0.70 ms per step).
//
// The step comes in two halves so that a caller with two batches can keep the device and the bus busy while the host
// patches: _begin enqueues everything (no host-side wait), _end waits chunk by chunk and patches the caller's arrays.
extern "C" int tg_step_host_sparse_begin(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                                         uint8_t *ran, void *stream) {
    if (!env || !actions || !obs || !reward || !done || !ran) return fail(TG_ERR_ARG, "null argument");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_step_host_sparse_begin: the previous step has not been ended");
    if (!env->sp_primed || !env->B.auto_reset || obs != env->sp_obs || reward != env->sp_reward || done != env->sp_done || ran != env->sp_ran) {
        // first call / other arrays: everything crosses, now.  Also without auto-reset: an env that stays done reports it in
        // every step without being touched, which the records do not carry.
        const int rc = tg_step_host(env, actions, obs, reward, done, ran, stream);
        if (rc == TG_OK) env->sp_pending = 2;
        return rc;
    }
    DeviceGuard guard(env->device);
    const double t_begin = now_s();
    int rc = ensure_staging(env);
    if (rc != TG_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = env->B.n;
    const int od = env->B.obs_dim;
    static const int forced_chunks = getenv("TG_SPARSE_CHUNKS") ? atoi(getenv("TG_SPARSE_CHUNKS")) : 0;
    // chunks: the kernel of chunk c + 1 and the copy of chunk c overlap the host-side patching of chunk c - 1 (measured at
    // 1,048,576 envs, one B200 + 16 host threads: 1 / 2 / 3 / 4 chunks = 0.68 / 0.59 / 0.48 / 0.53 ms per step)
    int chunks = 1;
    if (n >= (int64_t)6 * 131072) chunks = 3; else if (n >= (int64_t)3 * 131072) chunks = 2;
    if (forced_chunks >= 1 && forced_chunks <= 8) chunks = forced_chunks;
    const int64_t per = ((n + chunks - 1) / chunks + 2047) / 2048 * 2048;
    const int used = (int)((n + per - 1) / per);
    if (!env->sp_drecs) {
        env->sp_words = (3 + od + 3) / 4 * 4;                                   // 16-byte records
        // records + per chunk: a header and the tile table (two words per CTA; tiles hold >= 32 envs)
        const size_t total = ((size_t)n + 8) * env->sp_words + 8 * 4 + 2 * ((size_t)n / 32 + 8 * 4);
        CU(dev_alloc(env, &env->sp_drecs, total));
        CU(cudaHostAlloc((void **)&env->sp_hrecs, total * 4, cudaHostAllocDefault));
        for (int c = 0; c < 8; c++) CU(cudaEventCreateWithFlags(&env->ev_rec[c], cudaEventDisableTiming));
        for (int c = 0; c < 8; c++) CU(cudaEventCreateWithFlags(&env->ev_h2d[c], cudaEventDisableTiming));
        for (int c = 0; c < 8; c++) env->sp_guess[c] = -1;                      // first time: the whole region
    }
    if (!env->side) {
        CU(cudaStreamCreateWithFlags(&env->side, cudaStreamNonBlocking));
        for (int c = 0; c < 8; c++) CU(cudaEventCreateWithFlags(&env->ev_chunk[c], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&env->ev_join, cudaEventDisableTiming));
    }
    const int words = env->sp_words;
    // actions: chunk 0 on the caller's stream, the others on the side stream (its copy overlaps the kernel of chunk 0)
    for (int c = 0; c < used; c++) {
        const int64_t lo = (int64_t)c * per, cnt = (lo + per <= n) ? per : n - lo;
        CU(cudaMemcpyAsync(env->s_actions + lo, actions + lo, (size_t)cnt * sizeof(int32_t), cudaMemcpyHostToDevice, c == 0 ? s : env->side));
        if (c > 0) CU(cudaEventRecord(env->ev_h2d[c], env->side));
    }
    env->h2d_bytes += n * (int64_t)sizeof(int32_t);
    env->obs_sync = nullptr;                         // only records leave the device: no device buffer follows this step
    env->sp_table_off = 0;
    for (int c = 0; c < used; c++) {
        const int64_t lo = (int64_t)c * per, cnt = (lo + per <= n) ? per : n - lo;
        if (c > 0) CU(cudaStreamWaitEvent(s, env->ev_h2d[c], 0));
        const int tile = pick_step_tile(cnt, env->step_tile);
        const int grid = (int)((cnt + tile - 1) / tile);
        // region c = header c (16 bytes: the record count), the tile table of chunk c's launch, the records of chunk c
        const size_t off = (size_t)lo * words + (size_t)c * 4 + env->sp_table_off;
        env->sp_off[c] = off; env->sp_grid[c] = grid; env->sp_tile[c] = tile;
        const size_t tw = (2 * (size_t)grid + 3) & ~(size_t)3;                    // table words, records stay 16-byte aligned
        env->sp_table_off += tw;
        uint32_t *dreg = env->sp_drecs + off, *hreg = env->sp_hrecs + off;
        CU(cudaMemsetAsync(dreg, 0, 16, s));
        BatchView V = env->B;
        V.r_begin = lo; V.r_count = cnt; V.advance = (c == used - 1) ? 1 : 0;
        V.sp_count = dreg; V.sp_recs = dreg + 4 + tw; V.sp_words = words;
        CU(launch_step(V, env->ni, env->step_tile, env->s_actions, nullptr, nullptr, nullptr, nullptr, nullptr, s));
        env->launches++;
        CU(cudaEventRecord(env->ev_chunk[c], s));
        CU(cudaStreamWaitEvent(env->side, env->ev_chunk[c], 0));
        int64_t guess = env->sp_guess[c] < 0 ? cnt : env->sp_guess[c];
        if (guess > cnt) guess = cnt;
        env->sp_copied[c] = guess;
        CU(cudaMemcpyAsync(hreg, dreg, 16 + 4 * tw + (size_t)guess * words * 4, cudaMemcpyDeviceToHost, env->side));
        CU(cudaEventRecord(env->ev_rec[c], env->side));
    }
    CU(cudaEventRecord(env->ev_join, env->side));    // later work on the caller's stream comes after the record copies
    CU(cudaStreamWaitEvent(s, env->ev_join, 0));
    env->sp_pending = 1; env->sp_used = used; env->sp_per = per;
    env->sp_t_enqueue += now_s() - t_begin;
    return TG_OK;
}

extern "C" int tg_step_host_sparse_end(tg_env *env) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    if (!env->sp_pending) return fail(TG_ERR_STATE, "tg_step_host_sparse_end without a begin");
    if (env->sp_pending == 2) { env->sp_pending = 0; return TG_OK; }             // the dense path ran inside begin
    env->sp_pending = 0;
    DeviceGuard guard(env->device);
    const int64_t n = env->B.n, per = env->sp_per;
    const int od = env->B.obs_dim, words = env->sp_words, used = env->sp_used;
    float *obs = (float *)env->sp_obs, *reward = (float *)env->sp_reward;   // the arrays _begin was given (non-const there)
    uint8_t *done = (uint8_t *)env->sp_done, *ran = (uint8_t *)env->sp_ran;
    for (int c = 0; c < used; c++) {                 // patch chunk c while chunk c + 1 is still running / crossing
        const int64_t lo = (int64_t)c * per, cnt = (lo + per <= n) ? per : n - lo;
        const size_t off = env->sp_off[c];
        const int grid = env->sp_grid[c];
        uint32_t *dreg = env->sp_drecs + off, *hreg = env->sp_hrecs + off;
        const size_t tw = (2 * (size_t)grid + 3) & ~(size_t)3;
        uint32_t *drecs = dreg + 4 + tw, *hrecs = hreg + 4 + tw;
        const double t_w0 = now_s();
        CU(cudaEventSynchronize(env->ev_rec[c]));
        const double t_w1 = now_s();
        env->sp_t_wait += t_w1 - t_w0;
        const int64_t count = hreg[0];
        const int64_t copied = env->sp_copied[c];
        if (count > cnt) return fail(TG_ERR_STATE, "sparse step: %lld records for a chunk of %lld envs", (long long)count, (long long)cnt);
        if (count > copied) {                        // the guess was short (e.g. a step in which every env is reset): the rest
            CU(cudaMemcpyAsync(hrecs + (size_t)copied * words, drecs + (size_t)copied * words,
                               (size_t)(count - copied) * words * 4, cudaMemcpyDeviceToHost, env->side));
            CU(cudaStreamSynchronize(env->side));
        }
        env->d2h_bytes += 16 + 4 * (int64_t)tw + (copied > count ? copied : count) * words * 4;
        env->sp_guess[c] = count + count / 32 + (cnt / 256 > 64 ? cnt / 256 : 64);
        if (!sparse_apply_tiles(hreg + 4, grid, env->sp_tile[c], lo, cnt, hrecs, words, od, obs, reward, done, ran))
            return fail(TG_ERR_STATE, "sparse step: a record outside its tile");
        env->sp_t_patch += now_s() - t_w1;
    }
    // every chunk's record copy has completed, and each waited for its kernel: nothing of this step is left on the device
    return TG_OK;
}

extern "C" int tg_step_host_sparse(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                                   uint8_t *ran, void *stream) {
    const int rc = tg_step_host_sparse_begin(env, actions, obs, reward, done, ran, stream);
    return rc != TG_OK ? rc : tg_step_host_sparse_end(env);
}

extern "C" int tg_debug_level_closure(const tg_level *lv, int32_t obj, int32_t value, uint32_t bits, uint32_t *entry) {
    if (!lv || !entry) return fail(TG_ERR_ARG, "null argument");
    if (obj < 0 || obj >= lv->blob.n_objs || bits >= (1u << CLOSURE_BITS) || lv->closure.empty()) return fail(TG_ERR_ARG, "closure index outside the table");
    *entry = lv->closure[(((size_t)obj * 2 + (value ? 1 : 0)) << CLOSURE_BITS) | bits];
    return TG_OK;
}

extern "C" void tg_debug_host_times(tg_env *env, double *out3) {
    if (!env || !out3) return;
    out3[0] = env->sp_t_enqueue; out3[1] = env->sp_t_wait; out3[2] = env->sp_t_patch;
    env->sp_t_enqueue = env->sp_t_wait = env->sp_t_patch = 0;
}

extern "C" int tg_available_mask(tg_env *env, uint8_t *mask, void *stream) {
    if (!env || !mask) return fail(TG_ERR_ARG, "null argument");
    DeviceGuard guard(env->device);
    CU(launch_mask(env->B, env->ni, mask, (cudaStream_t)stream));
    env->launches++;
    return TG_OK;
}

extern "C" int tg_render(tg_env *env, int64_t first, int64_t count, uint8_t *frames, void *stream) {
    if (!env || !frames) return fail(TG_ERR_ARG, "null argument");
    if (!env->has_render) return fail(TG_ERR_STATE, "render needs tg_level_set_sprites on every level (and equal grid sizes)");
    if (first < 0 || count < 0 || first + count > env->B.n) return fail(TG_ERR_ARG, "frame range outside the batch");
    if (reinterpret_cast<uintptr_t>(frames) & 15u) return fail(TG_ERR_ARG, "frames must be 16-byte aligned (bulk stores)");
    if (count == 0) return TG_OK;
    DeviceGuard guard(env->device);
    bool have_lists = false;
    for (int l = 0; l < env->n_levels; l++) have_lists = have_lists || !env->lev_envs[l].empty();
    if (env->B.level_id && have_lists) {
        // mixed batch: the envs of [first, first + count) layout by layout
        const int32_t *lists[TG_MAX_LEVELS]; int64_t counts[TG_MAX_LEVELS];
        int nl = 0;
        for (int l = 0; l < env->n_levels; l++) {
            const std::vector<int32_t> &v = env->lev_envs[l];
            const size_t lo = std::lower_bound(v.begin(), v.end(), (int32_t)first) - v.begin();
            const size_t hi = std::lower_bound(v.begin(), v.end(), (int32_t)(first + count)) - v.begin();
            lists[l] = env->d_lev_envs[l] ? env->d_lev_envs[l] + lo : nullptr; counts[l] = (int64_t)(hi - lo);
            nl += counts[l] > 0;
        }
        CU(launch_render(env->B, env->R, first, count, frames, (cudaStream_t)stream, lists, counts));
        env->launches += nl;
        return TG_OK;
    }
    CU(launch_render(env->B, env->R, first, count, frames, (cudaStream_t)stream));
    env->launches += (count + 32767) / 32768;
    return TG_OK;
}

extern "C" int tg_step_frames(tg_env *env, const int64_t *env_ids, int32_t count, const int32_t *actions, int32_t max_ticks,
                              uint8_t *frames, int32_t *n_ticks, float *obs, float *reward, uint8_t *done, uint8_t *ran, void *stream) {
    if (!env || !env_ids || !actions || !frames || !n_ticks) return fail(TG_ERR_ARG, "null argument");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_step_frames between tg_step_host_sparse_begin and _end");
    if (!env->has_render) return fail(TG_ERR_STATE, "frames need tg_level_set_sprites on every level (and equal grid sizes)");
    if (count < 1 || max_ticks < 1 || (int64_t)count * max_ticks > (int64_t)1 << 24) return fail(TG_ERR_ARG, "count >= 1, max_ticks >= 1, count * max_ticks <= 2^24");
    if (reinterpret_cast<uintptr_t>(frames) & 15u) return fail(TG_ERR_ARG, "frames must be 16-byte aligned (bulk stores)");
    DeviceGuard guard(env->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t slots = (int64_t)count * max_ticks;
    if (slots > env->tr_slots) {                     // snapshot arrays: a batch of count * max_ticks states for the renderer
        CU(cudaStreamSynchronize(s));                // an earlier call's renderer may still read the old arrays
        for (void *p : {(void *)env->tr_core, (void *)env->tr_items, (void *)env->tr_angles, (void *)env->tr_lid}) if (p) cudaFree(p);
        env->tr_core = nullptr; env->tr_items = nullptr; env->tr_angles = nullptr; env->tr_lid = nullptr; env->tr_slots = 0;
        CU(cudaMalloc((void **)&env->tr_core, (size_t)slots * sizeof(uint4)));
        CU(cudaMalloc((void **)&env->tr_items, (size_t)slots * sizeof(uint2)));
        CU(cudaMalloc((void **)&env->tr_angles, (size_t)slots * TG_MAX_HANDLES * sizeof(double)));
        CU(cudaMalloc((void **)&env->tr_lid, (size_t)slots));
        env->tr_slots = slots;
    }
    BatchView S = env->B;
    S.n = slots; S.r_begin = 0; S.r_count = slots;
    S.core = env->tr_core; S.items23 = env->B.items23 ? env->tr_items : nullptr; S.angles = env->tr_angles;
    S.level_id = env->B.level_id ? env->tr_lid : nullptr;
    CU(launch_trace(env->B, S, env->ni, env_ids, count, actions, max_ticks, n_ticks, obs, reward, done, ran, s));
    CU(launch_render(S, env->R, 0, slots, frames, s));
    env->launches += 1 + (slots + 32767) / 32768;
    env->sp_primed = false;
    env->obs_sync = nullptr;
    return TG_OK;
}

extern "C" int tg_blend(tg_env *env, int64_t first, int64_t count, int64_t per_surface, uint8_t *surfaces,
                        int32_t alpha_objs, int32_t alpha_player, void *stream) {
    if (!env || !surfaces) return fail(TG_ERR_ARG, "null argument");
    if (!env->has_render) return fail(TG_ERR_STATE, "blend needs tg_level_set_sprites on every level (and equal grid sizes)");
    if (first < 0 || count < 0 || first + count > env->B.n) return fail(TG_ERR_ARG, "env range outside the batch");
    if (per_surface < 1 || count % per_surface != 0) return fail(TG_ERR_ARG, "count must be a multiple of per_surface");
    if (alpha_objs < 0 || alpha_objs > 255 || alpha_player < 0 || alpha_player > 255) return fail(TG_ERR_ARG, "opacities are 0..255");
    if (reinterpret_cast<uintptr_t>(surfaces) & 15u) return fail(TG_ERR_ARG, "surfaces must be 16-byte aligned");
    if (count == 0) return TG_OK;
    DeviceGuard guard(env->device);
    CU(launch_blend(env->B, env->R, first, count / per_surface, per_surface, surfaces, alpha_objs, alpha_player, (cudaStream_t)stream));
    env->launches += (count / per_surface + 32767) / 32768;
    return TG_OK;
}

extern "C" int tg_background(tg_env *env, int32_t level, uint8_t *frame, void *stream) {
    if (!env || !frame) return fail(TG_ERR_ARG, "null argument");
    if (!env->has_render) return fail(TG_ERR_STATE, "no render assets (tg_level_set_sprites)");
    if (level < 0 || level >= env->n_levels) return fail(TG_ERR_ARG, "level %d outside 0..%d", level, env->n_levels - 1);
    DeviceGuard guard(env->device);
    CU(cudaMemcpyAsync(frame, env->R.assets[level].background, (size_t)env->R.frame_w * env->R.frame_h * 3,
                       cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TG_OK;
}

extern "C" int tg_blit_alpha(uint8_t *target, int32_t tw, int32_t th, const uint8_t *source, int32_t sw, int32_t sh,
                             int32_t channels, int32_t x, int32_t y, int32_t opacity, void *stream) {
    if (!target || !source) return fail(TG_ERR_ARG, "null argument");
    if (tw < 1 || th < 1 || sw < 1 || sh < 1 || (channels != 3 && channels != 4)) return fail(TG_ERR_ARG, "bad surface geometry");
    if (opacity < 0 || opacity > 255) return fail(TG_ERR_ARG, "opacity is 0..255");
    CU(launch_blit_alpha(target, tw, th, source, sw, sh, channels, x, y, opacity, (cudaStream_t)stream));
    return TG_OK;
}

extern "C" int tg_get_state(tg_env *env, const tg_state_view *out, void *stream) {
    if (!env || !out) return fail(TG_ERR_ARG, "null argument");
    DeviceGuard guard(env->device);
    CU(launch_get_state(env->B, *out, (cudaStream_t)stream));
    env->launches++;
    return TG_OK;
}

extern "C" int tg_set_state(tg_env *env, const tg_state_view *in, void *stream) {
    if (!env || !in) return fail(TG_ERR_ARG, "null argument");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_set_state between tg_step_host_sparse_begin and _end");
    DeviceGuard guard(env->device);
    CU(launch_set_state(env->B, *in, (cudaStream_t)stream));
    env->launches++;
    env->sp_primed = false;
    env->obs_sync = nullptr;
    return TG_OK;
}

extern "C" int tg_primitive_step(tg_env *env, const int32_t *actions, float *obs, float *reward, uint8_t *done, void *stream) {
    if (!env || !actions) return fail(TG_ERR_ARG, "null argument");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_primitive_step between tg_step_host_sparse_begin and _end");
    DeviceGuard guard(env->device);
    CU(launch_primitive(env->B, env->ni, actions, obs, reward, done, (cudaStream_t)stream));
    env->launches++;
    env->sp_primed = false;
    env->obs_sync = obs;                 // every env's row is written
    return TG_OK;
}

extern "C" int tg_init_with_state(tg_env *env, const double *states, const uint8_t *mask, void *stream) {
    if (!env || !states) return fail(TG_ERR_ARG, "null argument");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_init_with_state between tg_step_host_sparse_begin and _end");
    DeviceGuard guard(env->device);
    CU(launch_init_with_state(env->B, env->ni, states, mask, (cudaStream_t)stream));
    env->launches++;
    env->sp_primed = false;
    env->obs_sync = nullptr;
    return TG_OK;
}

extern "C" int tg_set_draw_tape(tg_env *env, const double *tape, const int64_t *offsets, void *stream) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    if (env->sp_pending) return fail(TG_ERR_STATE, "tg_set_draw_tape between tg_step_host_sparse_begin and _end");
    if ((tape == nullptr) != (offsets == nullptr)) return fail(TG_ERR_ARG, "tape and offsets must both be set or both be null");
    DeviceGuard guard(env->device);
    env->B.tape = tape; env->B.tape_off = offsets;
    // draw indices restart at 0: acct.x is the first word of each 16-byte record
    CU(cudaMemset2DAsync(env->B.acct, sizeof(uint4), 0, sizeof(uint32_t), (size_t)env->B.n, (cudaStream_t)stream));
    return TG_OK;
}

extern "C" int tg_debug_phase_buffer(tg_env *env, uint64_t *stamps) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    env->B.phase_ts = reinterpret_cast<unsigned long long *>(stamps);
    return TG_OK;
}

extern "C" int tg_debug_set_step_tile(tg_env *env, int32_t tile) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    if (tile != 0 && (tile < 32 || tile > 4096)) return fail(TG_ERR_ARG, "tile must be 0 (automatic) or 32..4096");
    env->step_tile = tile;
    return TG_OK;
}

extern "C" int tg_stats(tg_env *env, int64_t *out8, void *stream) {
    if (!env || !out8) return fail(TG_ERR_ARG, "null argument");
    DeviceGuard guard(env->device);
    CU(cudaMemcpyAsync(out8, env->B.stats, 64, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TG_OK;
}

extern "C" int tg_stats_clear(tg_env *env, void *stream) {
    if (!env) return fail(TG_ERR_ARG, "null env");
    DeviceGuard guard(env->device);
    CU(cudaMemsetAsync(env->B.stats, 0, 64, (cudaStream_t)stream));
    return TG_OK;
}
