// Device-side dynamics of the Treasure Game: one thread owns one environment.
//
// The formulation is cell-granular bit-tests on a padded byte table in shared
// memory (the reference's 672x624 character map is provably cell-granular:
// _treasure_game_impl.py:204-216 appends the same row list 48 times and doors patch
// whole cells, _objects.py:246-253), state packed in registers, counter-based RNG.
// Citations: impl = _treasure_game_impl/_treasure_game_impl.py, opts = _move_options.py,
// opt = _option.py, objs = _objects.py, tg = treasure_game.py (all under
// /root/reference/gym_treasure_game/envs/).
#pragma once
#include <cuda_runtime.h>
#include <climits>
#include "tg_types.h"

namespace tg {

// ---------------------------------------------------------------------------
// per-thread working state
// ---------------------------------------------------------------------------
template <int NI>
struct Env {
    int px, py;
    uint32_t flags;
    uint32_t sticky;             // per handle: previously_triggered left raised by init_with_state (impl:473)
    int ix[NI], iy[NI];          // item (key / gold) pixel positions
    uint32_t draws;              // uniforms consumed by this env since creation
    uint32_t total_actions;      // impl:53,295 (per episode)
    // RNG
    uint32_t d0, w0, w1, w2, w3;    // draw index at the start of this API call, cached Philox block
    uint32_t key0, key1, id_lo, id_hi;
    const double *tape;             // parity mode: this env's slice of the draw tape
    uint32_t tape_len;              // ... and its length: a draw past the end flags the env (F_ERROR) and reads 0
    // row-mask cache for the probes every tick makes (valid while the probed pair of rows is the one in the key and the doors
    // do not move): m_fall = non-open cells of the rows of y and y+50, m_side = solid cells of the
    // rows of y+4 and y+44; bit = padded column.  A walk never changes y, so its ticks only shift/test.
    int m_py, m_py_side;         // key of the row pair each mask was built for: r_a | r_b << 8 | kind << 16 (INT_MIN = empty)
    uint32_t m_fall, m_side;
    // where this env's handle angles live (touched only by interact / reset / obs)
    double *angles;              // angles[h * n] is handle h
    int64_t n;
};

__device__ __forceinline__ bool facing(uint32_t f) { return (f >> F_FACING) & 1u; }
__device__ __forceinline__ int ticker(uint32_t f) { return (f >> F_TICKER) & 31u; }
__device__ __forceinline__ uint32_t set_ticker(uint32_t f, int t) {
    return (f & ~(31u << F_TICKER)) | ((uint32_t)t << F_TICKER);
}
__device__ __forceinline__ int bag_len(uint32_t f) { return (f >> F_BAGLEN) & 7u; }
// set of item indices present in the bag (impl:418-428 player_got_key / player_got_goldcoin)
__device__ __forceinline__ uint32_t bag_items(uint32_t f) {
    const int len = bag_len(f);
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < TG_MAX_ITEMS; j++) if (j < len) m |= 1u << ((f >> (F_BAGORD + 2 * j)) & 3u);
    return m;
}

// ---- packed state words -----------------------------------------------------
__device__ __forceinline__ uint32_t pack_xy(int x, int y) { return ((uint32_t)x & 0xFFFFu) | ((uint32_t)y << 16); }
// core.x = playerx (12 bits, 0 <= x < 26*48) | sticky handle flags (4 bits) | playery << 16
__device__ __forceinline__ uint32_t pack_player(int px, int py, uint32_t sticky) {
    return ((uint32_t)px & 0xFFFu) | ((sticky & 15u) << 12) | ((uint32_t)py << 16);
}
__device__ __forceinline__ int core_px(uint32_t v) { return (int)(v & 0xFFFu); }
__device__ __forceinline__ uint32_t core_sticky(uint32_t v) { return (v >> 12) & 15u; }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(int16_t)(v & 0xFFFFu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }

// ---------------------------------------------------------------------------
// RNG: Philox4x32-10, four 32-bit uniforms per block (u = word / 2^32).
// ctr = (d0, j >> 2, env_id_lo, env_id_hi), key = (seed_lo, seed_hi); word = j & 3 (see draw_k)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1,
                                              uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}
// Blocks (c0, b) and (c0, b + 1) side by side: two independent dependency chains for the fast-forward loop of the walkers
__device__ __forceinline__ void philox4x32_10_x2(uint32_t c0, uint32_t b, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                                 uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3,
                                                 uint32_t &y0, uint32_t &y1, uint32_t &y2, uint32_t &y3) {
    uint32_t a0 = c0, a1 = b, a2 = c2, a3 = c3, d0 = c0, d1 = b + 1u, d2 = c2, d3 = c3;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t ah0 = __umulhi(0xD2511F53u, a0), al0 = 0xD2511F53u * a0, ah1 = __umulhi(0xCD9E8D57u, a2), al1 = 0xCD9E8D57u * a2;
        const uint32_t dh0 = __umulhi(0xD2511F53u, d0), dl0 = 0xD2511F53u * d0, dh1 = __umulhi(0xCD9E8D57u, d2), dl1 = 0xCD9E8D57u * d2;
        a0 = ah1 ^ a1 ^ k0; a1 = al1; a2 = ah0 ^ a3 ^ k1; a3 = al0;
        d0 = dh1 ^ d1 ^ k0; d1 = dl1; d2 = dh0 ^ d3 ^ k1; d3 = dl0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    x0 = a0; x1 = a1; x2 = a2; x3 = a3; y0 = d0; y1 = d1; y2 = d2; y3 = d3;
}
// One copy of the block function per kernel, called from every draw site: inlined at each of the ~10 sites it was
// 14 KB of a 118 KB kernel body whose warps (different option classes side by side) kept missing the 32 KB
// instruction cache (profiles/r02_step_v13_ncu.txt: no-instruction 5.0 stalled warps per issue).
static __device__ __noinline__ uint4 philox_block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    uint4 o;
    philox4x32_10(c0, c1, c2, c3, k0, k1, o.x, o.y, o.z, o.w);
    return o;
}

// One uniform draw as the integer k with u = k / 2^53.  Tape mode: k is the recorded CPython double times
// 2^53 (exact: MT outputs are multiples of 2^-53).  Philox mode: one 32-bit word per draw, u = w / 2^32,
// i.e. k = w << 21, four draws per Philox4x32-10 block.  Blocks are numbered *per API call*: draw j of a call
// that began at draw index d0 is word (j & 3) of block (d0, j >> 2) -- so every lane of a warp refills on
// the same loop trip (j = 0, 4, 8 ...) however many draws its env consumed in earlier calls.  With blocks
// numbered by the absolute draw index the refills of 32 lanes were spread over all trips and the warp
// paid for the block function on nearly every tick (measured: 44 % of the time of a walking tick).
// (d0, j >> 2) never repeats for one env: a call that consumed draws advances d0 past them.
// Every transform of the reference is then applied to u exactly as CPython would (same thresholds, same
// round-half-even); only the resolution of u is 2^-32.
template <int NI>
__device__ __forceinline__ uint32_t draw_w(Env<NI> &e) {           // Philox mode: the 32-bit word of the next draw
    uint32_t j = e.draws++ - e.d0;
    if ((j & 3u) == 0u) {
        const uint4 o = philox_block(e.d0, j >> 2, e.id_lo, e.id_hi, e.key0, e.key1);
        e.w0 = o.x; e.w1 = o.y; e.w2 = o.z; e.w3 = o.w;
    }
    const uint32_t lo = (j & 1u) ? e.w1 : e.w0, hi = (j & 1u) ? e.w3 : e.w2;
    return (j & 2u) ? hi : lo;
}
template <bool TAPE, int NI>
__device__ __forceinline__ uint64_t draw_k(Env<NI> &e) {
    if (TAPE) {
        const uint32_t j = e.draws++;
        if (j >= e.tape_len) { e.flags |= 1u << F_ERROR; return 0ull; }     // tape exhausted: never read out of bounds
        return (uint64_t)__double2ull_rz(e.tape[j] * 9007199254740992.0);
    }
    return (uint64_t)draw_w(e) << 21;
}
__device__ __forceinline__ double k_to_unit(uint64_t k) { return (double)k * (1.0 / 9007199254740992.0); }   // exact
template <bool TAPE, int NI>
__device__ __forceinline__ double draw(Env<NI> &e) { return k_to_unit(draw_k<TAPE>(e)); }

// CPython random.uniform(a, b) = a + (b - a) * random(): two separately rounded operations.
__device__ __forceinline__ double uniform_span(double lo, double span, double u) {
    return __dadd_rn(lo, __dmul_rn(span, u));
}
// objs:111-114 / objs:127-131: angle for an up (0.85..1) or down (0..0.15) handle
__device__ __forceinline__ double handle_angle(bool up, double u) {
    // (1.0 - 0.85) evaluates to 0.15000000000000002 in binary64, as in CPython
    return up ? uniform_span(0.85, 1.0 - 0.85, u) : uniform_span(0.0, 0.15, u);
}

// ---------------------------------------------------------------------------
// tiles
// ---------------------------------------------------------------------------
// pixel -> padded cell index floor(v/48)+PAD.  No clamps: px in [0, W) and py in [-S, H) are invariants of
// the dynamics (tg_set_state / tg_init_with_state clamp injected positions), so every probe lands inside the
// 32x32 table.  t/48 == (t*1366)>>16 exactly for 0 <= t < 1536 = 32*48 (1366/65536 exceeds 1/48 by 1.0e-5,
// which adds < 0.016 to the quotient; the largest fractional part is 47/48 = 0.979), one IMAD + one shift.
static_assert(S == 48 && TSTRIDE == 32, "pad_cell's multiply-shift division is derived for 48-px cells and a 32-cell table");
__device__ __forceinline__ int pad_cell(int v) { return (int)(((unsigned)(v + PAD * S) * 1366u) >> 16); }
// v mod 48 for v >= -PAD*S
__device__ __forceinline__ int mod48(int v) { return v + PAD * S - S * pad_cell(v); }
__device__ __forceinline__ int pad_idx(int c) { return min(max(c + PAD, 0), TSTRIDE - 1); }

// effective type of the cell with padded indices (ixp, iyp)    impl:218-225 + objs:246-253
__device__ __forceinline__ int type_p(const LevelBlob &L, uint32_t flags, int ixp, int iyp) {
    int code = L.tiles[iyp * TSTRIDE + ixp];
    int d = (flags >> (F_DOORS + (code >> 4))) & 1u;
    return (code & TC_HAS_DOOR) ? d * T_DOOR : (code & 3);
}
// impl:227-230 object_type_at_cell
__device__ __forceinline__ int type_c(const LevelBlob &L, uint32_t flags, int cx, int cy) {
    return type_p(L, flags, pad_idx(cx), pad_idx(cy));
}

// impl:283-288  (x +- 10, y + {0, 50}) all OPEN
__device__ __forceinline__ bool can_fall(const LevelBlob &L, uint32_t flags, int px, int py) {
    int c0 = pad_cell(px - 10), c1 = pad_cell(px + 10), r0 = pad_cell(py), r1 = pad_cell(py + 50);
    int t = type_p(L, flags, c0, r0) | type_p(L, flags, c1, r0) | type_p(L, flags, c0, r1) | type_p(L, flags, c1, r1);
    return t == T_OPEN;
}
// impl:232-238  (x + {-4,0,4}, y + {-4..-1}) all OPEN; the probes span <= 2 columns and <= 2 rows
__device__ __forceinline__ bool up_clear(const LevelBlob &L, uint32_t flags, int px, int py) {
    int c0 = pad_cell(px - 4), c1 = pad_cell(px + 4), r0 = pad_cell(py - 4), r1 = pad_cell(py - 1);
    int t = type_p(L, flags, c0, r0) | type_p(L, flags, c1, r0) | type_p(L, flags, c0, r1) | type_p(L, flags, c1, r1);
    return t == T_OPEN;
}
// impl:240-250  y > 1 and a LADDER at (x +- 12, y + {-4, 0, 44})
__device__ __forceinline__ bool can_go_up(const LevelBlob &L, uint32_t flags, int px, int py) {
    if (py <= 1) return false;
    int c0 = pad_cell(px - 12), c1 = pad_cell(px + 12);
    int r0 = pad_cell(py - 4), r1 = pad_cell(py), r2 = pad_cell(py + 44);
    bool l = false;
    l |= type_p(L, flags, c0, r0) == T_LADDER; l |= type_p(L, flags, c1, r0) == T_LADDER;
    l |= type_p(L, flags, c0, r1) == T_LADDER; l |= type_p(L, flags, c1, r1) == T_LADDER;
    l |= type_p(L, flags, c0, r2) == T_LADDER; l |= type_p(L, flags, c1, r2) == T_LADDER;
    return l;
}
// impl:252-257  a LADDER at (x +- 12, y + 0..51): rows floor(y/48), +1, floor((y+51)/48)
__device__ __forceinline__ bool can_go_down(const LevelBlob &L, uint32_t flags, int px, int py) {
    int c0 = pad_cell(px - 12), c1 = pad_cell(px + 12);
    int r0 = pad_cell(py), r2 = pad_cell(py + 51), r1 = min(r0 + 1, r2);
    bool l = false;
    l |= type_p(L, flags, c0, r0) == T_LADDER; l |= type_p(L, flags, c1, r0) == T_LADDER;
    l |= type_p(L, flags, c0, r1) == T_LADDER; l |= type_p(L, flags, c1, r1) == T_LADDER;
    l |= type_p(L, flags, c0, r2) == T_LADDER; l |= type_p(L, flags, c1, r2) == T_LADDER;
    return l;
}
// impl:259-281  no WALL / DOOR at (x -+ 16, y + {4, 44}); WALL and DOOR are the odd type codes
__device__ __forceinline__ bool side_free(const LevelBlob &L, uint32_t flags, int x, int py) {
    int c = pad_cell(x);
    int t = type_p(L, flags, c, pad_cell(py + 4)) | type_p(L, flags, c, pad_cell(py + 44));
    return (t & 1) == 0;
}

// ---- row-mask forms of the per-tick probes ---------------------------------------------------
// closed doors of padded row r as column bits (objs:246-253: a closed door cell reads DOOR, an open one OPEN)
__device__ __forceinline__ uint32_t door_bits(const LevelBlob &L, uint32_t flags, int r) {
    const int li = L.row_lut[r];
    const uint32_t closed = (flags >> F_DOORS) & 63u;
    return li ? L.door_lut[li - 1][closed] : 0u;
}
// A mask depends on playery only through the two padded rows its probes fall in, so the cache is keyed by the row
// pair: a fall moves one pixel at a time and re-probes after every pixel (impl:341-346) -- 47 of 48 are hits.
// The fall slot also serves up_clear (impl:232-238; same cells kind, other rows: kind bit in the key) -- a jump's
// rise evaluates up_clear every tick and never can_fall, its descent the other way round.
__device__ __forceinline__ int fall_key(int py) { return pad_cell(py) | (pad_cell(py + 50) << 8); }
__device__ __forceinline__ int upc_key(int py) { return pad_cell(py - 4) | (pad_cell(py - 1) << 8) | (1 << 16); }
__device__ __forceinline__ int side_key(int py) { return pad_cell(py + 4) | (pad_cell(py + 44) << 8); }
template <int NI>
__device__ __forceinline__ void fall_cache_fill(Env<NI> &e, const LevelBlob &L, int key) {
    const int r0 = key & 255, r1 = (key >> 8) & 255;
    e.m_fall = L.row_nonopen[r0] | door_bits(L, e.flags, r0) | L.row_nonopen[r1] | door_bits(L, e.flags, r1);
    e.m_py = key;
}
template <int NI>
__device__ __forceinline__ void side_cache_fill(Env<NI> &e, const LevelBlob &L, int key) {
    const int r2 = key & 255, r3 = (key >> 8) & 255;
    e.m_side = L.row_solid[r2] | door_bits(L, e.flags, r2) | L.row_solid[r3] | door_bits(L, e.flags, r3);
    e.m_py_side = key;
}
template <int NI>
__device__ __forceinline__ void row_cache_drop(Env<NI> &e) { e.m_py = INT_MIN; e.m_py_side = INT_MIN; }
// impl:283-288 through the cache
template <int NI>
__device__ __forceinline__ bool can_fall_m(Env<NI> &e, const LevelBlob &L) {
    const int key = fall_key(e.py);
    if (e.m_py != key) fall_cache_fill(e, L, key);
    return (((e.m_fall >> pad_cell(e.px - 10)) | (e.m_fall >> pad_cell(e.px + 10))) & 1u) == 0u;
}
// impl:232-238 through the cache: (x + {-4,0,4}, y + {-4..-1}) all OPEN; the probes span <= 2 columns and <= 2 rows
template <int NI>
__device__ __forceinline__ bool up_clear_m(Env<NI> &e, const LevelBlob &L) {
    const int key = upc_key(e.py);
    if (e.m_py != key) fall_cache_fill(e, L, key);
    return (((e.m_fall >> pad_cell(e.px - 4)) | (e.m_fall >> pad_cell(e.px + 4))) & 1u) == 0u;
}
// impl:259-281 through the cache (x = playerx -+ 16)
template <int NI>
__device__ __forceinline__ bool side_free_m(Env<NI> &e, const LevelBlob &L, int x) {
    const int key = side_key(e.py);
    if (e.m_py_side != key) side_cache_fill(e, L, key);
    return ((e.m_side >> pad_cell(x)) & 1u) == 0u;
}
// impl:240-257 in mask form: up: y > 1 and LADDER in rows of y-4, y, y+44; down: rows of y, next, y+51; columns of x-+12
__device__ __forceinline__ bool ladder_probe(const LevelBlob &L, int px, int py, bool up) {
    const int r0 = pad_cell(up ? py - 4 : py), r1 = up ? pad_cell(py) : min(pad_cell(py) + 1, pad_cell(py + 51)),
              r2 = pad_cell(up ? py + 44 : py + 51);
    const uint32_t m = L.row_ladder[r0] | L.row_ladder[r1] | L.row_ladder[r2];      // door cells never read LADDER
    const bool hit = (((m >> pad_cell(px - 12)) | (m >> pad_cell(px + 12))) & 1u) != 0u;
    return hit && (!up || py > 1);
}

// objs:46-53 evaluated at (px, py + 24): integer form of sqrt(dx^2 + dy^2) < r
__device__ __forceinline__ bool near_px(int px, int py, int ox, int oy, int r2) {
    int dx = px - (ox + S / 2), dy = py - oy;
    return dx * dx + dy * dy < r2;
}

__device__ __forceinline__ bool has_key(const LevelBlob &L, uint32_t f) { return (bag_items(f) & L.key_mask) != 0; }
__device__ __forceinline__ bool has_gold(const LevelBlob &L, uint32_t f) { return (bag_items(f) & L.gold_mask) != 0; }

// ---------------------------------------------------------------------------
// trigger graph + INTERACT (rare: one tick per interact option) -- kept out of line
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool obj_value(const LevelBlob &L, uint32_t f, int o) {
    int k = L.obj_kind[o], i = L.obj_idx[o];
    int bit = (k == TG_DOOR) ? F_DOORS + i : (k == TG_HANDLE) ? F_HANDLES + i : F_BOLTS + i;
    return (f >> bit) & 1u;
}

// set_val of door / handle / bolt (objs:145-149, :175-178, :231-235); returns true when the value
// changed (the caller then runs process_trigger).  A handle redraws its angle (objs:127-131).
template <bool TAPE, int NI>
__device__ __forceinline__ bool apply_val(Env<NI> &e, const LevelBlob &L, int o, bool v) {
    int k = L.obj_kind[o], i = L.obj_idx[o];
    int bit = (k == TG_DOOR) ? F_DOORS + i : (k == TG_HANDLE) ? F_HANDLES + i : F_BOLTS + i;
    if ((bool)((e.flags >> bit) & 1u) == v) return false;
    e.flags ^= 1u << bit;
    if (k == TG_HANDLE) e.angles[(int64_t)i * e.n] = handle_angle(v, draw<TAPE>(e));
    return true;
}

// process_trigger (objs:76-94) as an explicit DFS from object o0 that has just taken value v0.  `pt` is the
// set of objects whose previously_triggered flag is raised: those on the DFS stack plus the handles that
// init_with_state left flagged (impl:473, Env::sticky); a flag is lowered when its object's own
// process_trigger returns (objs:94), which is also how a sticky flag eventually clears.
// The stack lives in registers: an object is on it at most once (its flag in `pt` blocks re-entry), so a level is the
// object (4 bits; its value is the one it holds now) and the position in its trigger list (8 bits).
template <bool TAPE, int NI>
__device__ __forceinline__ void trigger_dfs(Env<NI> &e, const LevelBlob &L, int o0, bool v0) {
    static_assert(TG_MAX_OBJECTS <= 16 && TG_MAX_TRIGGERS < 256, "trigger_dfs packs its stack into three 64-bit words");
    uint64_t st_obj = 0, st_t0 = 0, st_t1 = 0;                    // level l: object = nibble l, list position = byte l (st_t0: levels 0-7)
    int sp = 0;
    uint32_t pt = 0;
    for (int h = 0; h < L.n_handles; h++) if ((e.sticky >> h) & 1u) pt |= 1u << L.handle_obj[h];
    pt |= 1u << o0;
    auto set_t = [&](int l, uint32_t t) {
        const int sh = 8 * (l & 7);
        if (l < 8) st_t0 = (st_t0 & ~(0xFFull << sh)) | ((uint64_t)t << sh); else st_t1 = (st_t1 & ~(0xFFull << sh)) | ((uint64_t)t << sh);
    };
    st_obj = (uint64_t)o0; set_t(0, L.trig_begin[2 * o0 + (v0 ? 1 : 0)]); sp = 1;
    while (sp > 0) {
        const int l = sp - 1;
        const int o = (int)((st_obj >> (4 * l)) & 15ull);
        uint32_t t = (uint32_t)(((l < 8 ? st_t0 : st_t1) >> (8 * (l & 7))) & 0xFFull);
        const uint32_t tend = L.trig_begin[2 * o + (obj_value(L, e.flags, o) ? 1 : 0) + 1];
        bool pushed = false;
        while (t < tend) {
            const int ent = L.trig_list[t++];
            const int dst = ent & 127; const bool dv = ent >> 7;
            if (pt & (1u << dst)) continue;                       // objs:84 / :91
            if (apply_val<TAPE>(e, L, dst, dv)) {
                set_t(l, t);
                pt |= 1u << dst;
                st_obj = (st_obj & ~(15ull << (4 * sp))) | ((uint64_t)dst << (4 * sp));
                set_t(sp, L.trig_begin[2 * dst + (dv ? 1 : 0)]);
                sp++;
                pushed = true;
                break;
            }
        }
        if (!pushed) { pt &= ~(1u << o); sp--; }                   // objs:94
    }
    uint32_t sticky = 0;
    for (int h = 0; h < L.n_handles; h++) if ((pt >> L.handle_obj[h]) & 1u) sticky |= 1u << h;
    e.sticky = sticky;
}

// set_val (objs:145-149, :175-178, :231-235): nothing happens unless the value changes
// Without sticky handles the whole walk is a function of (o0, v0, door / handle / bolt bits): it is tabulated per level on the
// host (LevelBlob::closure, tg_capi.cu build_closure), and what is left here is one load, the new bits and the angle draws
// of the handles that moved, in the walk's order.  The walk itself stays for envs that init_with_state left with sticky
// handles, for levels whose cascades move more than five handles, and for batches created without the table.
template <bool TAPE, int NI>
__device__ __forceinline__ void set_val(Env<NI> &e, const LevelBlob &L, int o0, bool v0) {
    if (L.closure != nullptr && e.sticky == 0u) {
        const uint32_t ent = __ldg(L.closure + ((((uint32_t)o0 * 2u + (v0 ? 1u : 0u)) << CLOSURE_BITS) | ((e.flags >> F_DOORS) & ((1u << CLOSURE_BITS) - 1u))));
        const uint32_t ne = (ent >> CLOSURE_BITS) & 7u;
        if (ne != 7u) {
            e.flags = (e.flags & ~(((1u << CLOSURE_BITS) - 1u) << F_DOORS)) | ((ent & ((1u << CLOSURE_BITS) - 1u)) << F_DOORS);
            for (uint32_t k = 0; k < ne; k++) {
                const uint32_t ev = ent >> (16u + 3u * k);
                e.angles[(int64_t)(ev & 3u) * e.n] = handle_angle(((ev >> 2) & 1u) != 0u, draw<TAPE>(e));
            }
            return;
        }
    }
    if (apply_val<TAPE>(e, L, o0, v0)) trigger_dfs<TAPE>(e, L, o0, v0);
}

// impl:434-439: the first key in bag order leaves the bag and goes to cell (-1,-1)
template <int NI>
__device__ __forceinline__ void drop_key(Env<NI> &e, const LevelBlob &L) {
    int len = bag_len(e.flags);
    uint32_t ord = e.flags >> F_BAGORD;
    for (int j = 0; j < len; j++) {
        int it = (ord >> (2 * j)) & 3;
        if ((L.key_mask >> it) & 1) {
            uint32_t low = ord & ((1u << (2 * j)) - 1u);
            uint32_t high = (ord >> (2 * j + 2)) << (2 * j);
            ord = (low | high) & 0xFFu;
            e.flags = (e.flags & ~(0xFFu << F_BAGORD)) | (ord << F_BAGORD);
            e.flags = (e.flags & ~(7u << F_BAGLEN)) | ((uint32_t)(len - 1) << F_BAGLEN);   // player_bag.remove(obj)
#pragma unroll
            for (int q = 0; q < NI; q++) if (q == it) { e.ix[q] = -S; e.iy[q] = -S; }
            return;
        }
    }
}

// impl:321-329.  One call site of set_val for both kinds of object (the trigger walk is the bulk of this code).
template <bool TAPE, int NI>
__device__ __forceinline__ void interact(Env<NI> &e, const LevelBlob &L) {
    for (int o = 0; o < L.n_objs; o++) {
        const int k = L.obj_kind[o], i = L.obj_idx[o];
        bool set = false, val = false, drop = false;
        if (k == TG_HANDLE) {
            if (!near_px(e.px, e.py, L.handle_cx[i] * S, L.handle_cy[i] * S, 36 * 36)) continue;   // objs:115
            const bool up = (e.flags >> (F_HANDLES + i)) & 1u;
            if (draw_k<TAPE>(e) <= 7205759403792794ull) { set = true; val = !up; }   // objs:117-122: uniform(0,1) <= 0.8  (0.8 * 2^53)
            else e.angles[(int64_t)i * e.n] = handle_angle(up, draw<TAPE>(e));
        } else if (k == TG_BOLT) {
            if (!near_px(e.px, e.py, L.bolt_cx[i] * S, L.bolt_cy[i] * S, 24 * 24)) continue;
            if (has_key(L, e.flags)) { set = true; val = false; drop = true; }                    // impl:326-329
        }
        if (set) set_val<TAPE>(e, L, o, val);
        if (drop) drop_key(e, L);
    }
}

// ---------------------------------------------------------------------------
// primitive tick  (impl:290-359)
// ---------------------------------------------------------------------------
// impl:361-366: int(round(uniform(-4,-2))) or int(round(uniform(2,4))) with Python's round-half-even,
// evaluated exactly in integers.  With u = k/2^53: fl(c + 2u) = c + m*2^-51 where m = k/2 rounded to
// nearest-even (the binade [2,4) has ulp 2^-51), and rint(c + t) = c + rint(t) for the even integers
// c = -4, 2; rint(m/2^51) is 0 up to and including the tie m = 2^50, 2 from the tie m = 3*2^50 on.
// (Checked against CPython on 2e6 random and all threshold-adjacent k: tests/test_rng_integer_forms.py.)
__device__ __forceinline__ int noisy_from_k(uint64_t k, bool negative) {
    const uint64_t q = k >> 1, m = q + (k & q & 1ull);
    const int r = (int)(m > (1ull << 50)) + (int)(m >= (3ull << 50));
    return (negative ? -4 : 2) + r;
}
// For k = w << 21 (Philox mode) the same function in 32-bit arithmetic: q = w << 20, k & q & 1 = 0, m = w << 20.
__device__ __forceinline__ int noisy_r_from_w(uint32_t w) { return (int)(w > (1u << 30)) + (int)(w >= (3u << 30)); }
// The cached block holds the word of the next draw (refilled at a block boundary): for the straight-line loops, which
// then take their words with noisy_cached / consume_block.
template <int NI>
__device__ __forceinline__ void ensure_block(Env<NI> &e) {
    const uint32_t j = e.draws - e.d0;
    if ((j & 3u) == 0u) {
        const uint4 o = philox_block(e.d0, j >> 2, e.id_lo, e.id_hi, e.key0, e.key1);
        e.w0 = o.x; e.w1 = o.y; e.w2 = o.z; e.w3 = o.w;
    }
}
// Up to `limit` unchecked ticks of noisy(-4) (neg) / noisy(+4) from the words left in the cached block, branch-free.
// A tick may run when it starts inside the safe stretch, i.e. when the distance moved before it is <= D; the moves
// are positive (2, 3 or 4 px), so these are comparisons of the prefix sums with D.  Returns the number of ticks (>= 1
// when D >= 0), `moved` = pixels moved (absolute value); advances the draw index.
template <int NI>
__device__ __forceinline__ int consume_block(Env<NI> &e, bool neg, int D, int limit, int &moved) {
    const uint32_t a = (e.draws - e.d0) & 3u;
    uint32_t w0 = e.w0, w1 = e.w1, w2 = e.w2, w3 = e.w3;
    if (a & 1u) { w0 = w1; w1 = w2; w2 = w3; }
    if (a & 2u) { w0 = w2; w1 = w3; }
    const int avail = min(4 - (int)a, limit);
    const int base = neg ? 4 : 2, sg = neg ? -1 : 1;                               // |noisy(-4)| = 4 - r, noisy(+4) = 2 + r
    const int p1 = base + sg * noisy_r_from_w(w0), p2 = p1 + base + sg * noisy_r_from_w(w1),
              p3 = p2 + base + sg * noisy_r_from_w(w2), p4 = p3 + base + sg * noisy_r_from_w(w3);
    const bool ok0 = avail > 0 && D >= 0, ok1 = ok0 && avail > 1 && p1 <= D, ok2 = ok1 && avail > 2 && p2 <= D,
               ok3 = ok2 && avail > 3 && p3 <= D;
    const int u = (int)ok0 + (int)ok1 + (int)ok2 + (int)ok3;
    moved = ok3 ? p4 : ok2 ? p3 : ok1 ? p2 : ok0 ? p1 : 0;
    e.draws += (uint32_t)u;
    return u;
}

// noisy() for a caller that has just refilled the block at a block boundary (the straight-line option loops): the
// word of the next draw comes from the cached block, no refill check
template <bool TAPE, int NI>
__device__ __forceinline__ int noisy_cached(Env<NI> &e, bool negative) {
    if (TAPE) return noisy_from_k(draw_k<true>(e), negative);
    const uint32_t j = e.draws++ - e.d0;
    const uint32_t lo = (j & 1u) ? e.w1 : e.w0, hi = (j & 1u) ? e.w3 : e.w2;
    return (negative ? -4 : 2) + noisy_r_from_w((j & 2u) ? hi : lo);
}
template <bool TAPE, int NI>
__device__ __forceinline__ int noisy(Env<NI> &e, bool negative) {
    if (TAPE) return noisy_from_k(draw_k<true>(e), negative);
    return (negative ? -4 : 2) + noisy_r_from_w(draw_w(e));
}

// impl:350-354: key / gold within 24 px of (playerx, playery + 24) go to the next bag cell, item (= file) order
template <int NI>
__device__ __forceinline__ void pickups(Env<NI> &e, const LevelBlob &L) {
    const int bx = (L.cw - 1) * S, by = (L.ch - 1) * S;
#pragma unroll
    for (int i = 0; i < NI; i++) {
        if (i < L.n_items && near_px(e.px, e.py, e.ix[i], e.iy[i], 24 * 24)) {
            const int len = bag_len(e.flags);
            e.ix[i] = bx - len * S; e.iy[i] = by;                                    // objs:34-38
            if (len < TG_MAX_ITEMS) {                                                // player_bag.append(obj)
                e.flags = (e.flags & ~(3u << (F_BAGORD + 2 * len))) | ((uint32_t)i << (F_BAGORD + 2 * len));
                e.flags = (e.flags & ~(7u << F_BAGLEN)) | ((uint32_t)(len + 1) << F_BAGLEN);
            } else {
                e.flags |= 1u << F_ERROR;     // a fifth bag entry (only reachable through repeated re-pickups) is not representable
            }
        }
    }
}

// `ladder_ok` >= 0 passes the ladder probe the option policy has just evaluated for UP / DOWN (same state, same
// result as impl:298 / :302 would compute again); -1 = evaluate here.
// INTERACT_OK = false leaves out the INTERACT branch (trigger graph, handle angles: the step kernel runs the interact
// option out of line, interact_option_mem below, to keep that code out of its hot instruction footprint).
template <bool TAPE, int NI, bool INTERACT_OK = true>
__device__ __forceinline__ void tick(Env<NI> &e, const LevelBlob &L, int act, int ladder_ok = -1) {
    int xd = 0, yd = 0;
    e.total_actions++;
    if (act >= A_UP && act <= A_RIGHT) {
        // impl:297-313.  One instruction stream for the four moves (lanes of a warp mix them):
        // the move is permitted by the ladder probe (UP/DOWN) or the side probe (LEFT/RIGHT), then
        // noisy() draws once; LEFT/RIGHT also set facing_right (unchanged when blocked).
        const bool horiz = act >= A_LEFT, neg = (act == A_UP) || (act == A_LEFT);
        const bool ok = horiz ? side_free_m(e, L, e.px + (neg ? -16 : 16))
                              : (ladder_ok >= 0 ? ladder_ok != 0 : ladder_probe(L, e.px, e.py, neg));
        if (ok) {
            const int d = noisy<TAPE>(e, neg);
            if (horiz) { xd = d; e.flags = (e.flags & ~(1u << F_FACING)) | (neg ? 0u : (1u << F_FACING)); }
            else yd = d;
        }
    } else if (act == A_JUMP) {
        if (!ladder_probe(L, e.px, e.py, false) && up_clear_m(e, L))
            e.flags = set_ticker(e.flags, draw_k<TAPE>(e) > (1ull << 51) ? 23 : 22);   // impl:316-319: random() > 0.25
    } else if (INTERACT_OK && act == A_INTERACT) {
        interact<TAPE>(e, L);
        row_cache_drop(e);           // doors may have moved
    }
    int tk = ticker(e.flags);
    if (tk > 0) {                                                                    // impl:331-334
        if (up_clear_m(e, L)) yd = -4;
        e.flags = set_ticker(e.flags, tk - 1);
    } else if (can_fall_m(e, L)) {                                                   // impl:335-337
        yd = 4;
    }
    e.px += xd;                                                                      // impl:339
    if (yd > 0 && can_fall_m(e, L)) {                                                // impl:341-346
        // The reference moves one pixel at a time and re-probes can_fall after each.  The probe only depends on the
        // rows of y and y + 50: while y mod 48 stays below 46 they do not change, so every re-probe repeats the one
        // above and the whole move goes through.
        const int q = e.py + PAD * S - S * pad_cell(e.py);                           // y mod 48 (pad_cell = floor(y / 48) + PAD)
        if (q + yd <= 45) e.py += yd;
        else do {
            e.py++; yd--;
            if (!can_fall_m(e, L)) yd = 0;
        } while (yd > 0);
    } else {
        e.py += yd;                                                                  // impl:348
    }
    pickups(e, L);
}

// ---------------------------------------------------------------------------
// option layer  (opts + opt:20-36)
// ---------------------------------------------------------------------------
__device__ __forceinline__ int floordiv48(int v) { return (v >= 0) ? v / S : -((-v + S - 1) / S); }

template <int NI>
__device__ __forceinline__ void player_cell(const Env<NI> &e, int &cx, int &cy) {   // impl:441-445
    cx = pad_cell(e.px) - PAD; cy = pad_cell(e.py + S / 2) - PAD;     // exact floor: both arguments are >= -PAD*S
}

// impl:402-409 is_object_at: handle / bolt / key / gold in the cell, or a *closed* door
template <int NI>
__device__ __forceinline__ bool object_at(const Env<NI> &e, const LevelBlob &L, int cx, int cy) {
    int code = L.tiles[pad_idx(cy) * TSTRIDE + pad_idx(cx)];
    bool r = (code & TC_STATIC_OBJ) != 0;
    r |= (code & TC_HAS_DOOR) && ((e.flags >> (F_DOORS + (code >> 4))) & 1u);
#pragma unroll
    for (int i = 0; i < NI; i++)     // obj.cx/cy follow x/y (objs:34-44): trunc(x/48)
        r |= (i < L.n_items) && (e.ix[i] / S == cx) && (e.iy[i] / S == cy);   // C '/' truncates like int(x / 48)
    return r;
}

__device__ __forceinline__ bool closed_door_at(const LevelBlob &L, uint32_t flags, int cx, int cy) {  // impl:411-416
    int code = L.tiles[pad_idx(cy) * TSTRIDE + pad_idx(cx)];
    return (code & TC_HAS_DOOR) && ((flags >> (F_DOORS + (code >> 4))) & 1u);
}

// go_left (s=-1) / go_right (s=+1): can_run + target column (opts:23-67 / opts:95-139) on row bit masks.
// is_target_cell(xc) = ladder above/below | wall at xc+s | object at xc (handle, bolt, key, gold, closed
// door) | closed door at xc+s | open cell below xc+s; the target is the first such column beyond the
// player's; can_run = every cell from the player's to the target is OPEN over a non-OPEN cell.
template <int NI>
__device__ __forceinline__ bool walk_setup(const Env<NI> &e, const LevelBlob &L, int s, int &tcx) {
    const uint32_t f = e.flags;
    const int cxp = pad_cell(e.px), r = pad_cell(e.py + S / 2);          // padded player cell
    const uint32_t D = door_bits(L, f, r), D1 = door_bits(L, f, r + 1);  // closed doors of this row / the row below
    const uint32_t open_r = ~(L.row_nonopen[r] | D), open_r1 = ~(L.row_nonopen[r + 1] | D1);
    uint32_t obj = L.row_static_obj[r] | D;                              // impl:402-409
#pragma unroll
    for (int i = 0; i < NI; i++) {       // obj.cx/cy follow x/y (objs:34-44): C '/' truncates like int(x / 48)
        const int icx = e.ix[i] / S + PAD, icy = e.iy[i] / S + PAD;
        if (i < L.n_items && icy == r && icx >= 0 && icx < TSTRIDE) obj |= 1u << icx;
    }
    const uint32_t nb = L.row_solid[r] | D | open_r1;                    // the three tests that look at column xc + s
    const uint32_t tm = L.row_ladder[r - 1] | L.row_ladder[r + 1] | obj | (s < 0 ? nb << 1 : nb >> 1);
    // candidates strictly beyond the player's column, inside the grid on the left (opts:49-50: xc < 0 -> None);
    // on the right the border wall always yields a target
    const uint32_t cand = s < 0 ? tm & ((1u << cxp) - 1u) & ~((1u << PAD) - 1u) : tm & ~((2u << cxp) - 1u);
    if (!cand) return false;
    const int txp = s < 0 ? 31 - __clz(cand) : __ffs(cand) - 1;
    const uint32_t okm = open_r & ~open_r1;                              // opts:34-39
    const int lo = min(cxp, txp), hi = max(cxp, txp);
    const uint32_t range = ((2u << hi) - 1u) & ~((1u << lo) - 1u);
    tcx = txp - PAD;
    return (okm & range) == range;
}

__device__ __forceinline__ bool landing(const LevelBlob &L, uint32_t f, int cx, int cy) {   // opts:281-287
    return type_c(L, f, cx, cy) == T_OPEN && type_c(L, f, cx, cy + 1) == T_WALL;
}

// Evaluates can_run of option k (tg:83-89 / opt:22) and, for the options that walk to a
// column, the target column.  err is set when the reference would raise (target None).
template <int NI>
__device__ __forceinline__ bool option_setup(const Env<NI> &e, const LevelBlob &L, int k, int &tcx, bool &err) {
    const uint32_t f = e.flags;
    tcx = 0; err = false;
    int pcx, pcy;
    switch (k) {
    case TG_GO_LEFT:  return walk_setup(e, L, -1, tcx);
    case TG_GO_RIGHT: return walk_setup(e, L, +1, tcx);
    case TG_UP_LADDER:   return ladder_probe(L, e.px, e.py, true);    // opts:165-166 (impl:240-250)
    case TG_DOWN_LADDER: return ladder_probe(L, e.px, e.py, false);   // opts:181-182 (impl:252-257)
    case TG_INTERACT: {                                               // opts:446-455
        bool r = false;
        for (int i = 0; i < L.n_handles; i++)
            r |= near_px(e.px, e.py, L.handle_cx[i] * S, L.handle_cy[i] * S, 36 * 36);
        if (has_key(L, f))
            for (int i = 0; i < L.n_bolts; i++)
                r |= near_px(e.px, e.py, L.bolt_cx[i] * S, L.bolt_cy[i] * S, 24 * 24);
        return r;
    }
    case TG_DOWN_LEFT: case TG_DOWN_RIGHT: {                          // opts:199-221 / 394-416
        int s = (k == TG_DOWN_LEFT) ? -1 : 1;
        player_cell(e, pcx, pcy);
        if (type_c(L, f, pcx + s, pcy) != T_OPEN || type_c(L, f, pcx + s, pcy + 1) != T_OPEN) return false;
        tcx = pcx + s;
        int yc = pcy + 1;
        while (type_c(L, f, tcx, yc) == T_OPEN) { yc++; if (yc >= L.ch) { err = true; return false; } }
        return true;
    }
    case TG_JUMP_LEFT: case TG_JUMP_RIGHT: {                          // opts:254-279 / 324-349
        int s = (k == TG_JUMP_LEFT) ? -1 : 1;
        player_cell(e, pcx, pcy);
        if (type_c(L, f, pcx, pcy - 1) != T_OPEN || type_c(L, f, pcx + s, pcy - 1) != T_OPEN) return false;
        if (landing(L, f, pcx + s, pcy - 1)) { tcx = pcx + s; return true; }
        if (landing(L, f, pcx + 2 * s, pcy - 1)) { tcx = pcx + 2 * s; return true; }
        return false;
    }
    }
    return false;      // out-of-range action: the reference raises IndexError (tg:92); we report "not run"
}

// Conservative extent of the "nothing can happen" stretch of a go_left / go_right walk (ticker == 0, row masks
// mside / mfall valid for e.py).  Returns the farthest x in direction s such that at every player x' between
// e.px and x (inclusive): the ground holds (impl:283-288 false), the side probe is free (impl:259-281), the
// option's target is not yet reached (opts:69-72) and no key / gold is within pick-up range (impl:350-354).
// If e.px itself fails one of these the result lies behind the player (e.px - s).
// Cell c of the padded grid covers pixels [48c - 144, 48c - 97] (pad_cell).
template <int NI>
__device__ __forceinline__ int walk_safe_bound(const Env<NI> &e, const LevelBlob &L, int s, int tpx,
                                               uint32_t mside, uint32_t mfall) {
    const int px = e.px, behind = px - s, FAR = 1 << 20;
    const int a = pad_cell(px - 10), b = pad_cell(px + 10);
    if ((((mfall >> a) | (mfall >> b)) & 1u) == 0u) return behind;            // would fall now
    if (abs(tpx - px) < 4) return behind;                                        // aligned now
    int bound;
    if (s > 0) {
        bound = (px <= tpx - 4) ? tpx - 4 : FAR;
        const int c0 = pad_cell(px + 16);
        const uint32_t blk = mside >> c0;                                         // first blocking cell at / right of c0
        if (blk) bound = min(bound, 48 * (c0 + __ffs(blk) - 1) - 144 - 16 - 1);
        const uint32_t hole = ~mfall >> (a + 1);                                  // both probes over open cells: x - 10 enters it
        if (hole) bound = min(bound, 48 * (a + __ffs(hole)) - 144 + 10 - 1);
    } else {
        bound = (px >= tpx + 4) ? tpx + 4 : -FAR;
        const int c0 = pad_cell(px - 16);
        const uint32_t blk = mside << (31 - c0);                                  // bit c0 -> bit 31
        if (blk) bound = max(bound, 48 * (c0 - __clz(blk)) - 144 + 47 + 16 + 1);
        const uint32_t hole = ~mfall << (32 - b);                                 // bit b-1 -> bit 31 (b >= 3)
        if (hole) bound = max(bound, 48 * (b - 1 - __clz(hole)) - 144 + 47 - 10 + 1);
    }
#pragma unroll
    for (int i = 0; i < NI; i++) {
        if (i >= L.n_items) continue;
        const int dy = e.py - e.iy[i], c = e.ix[i] + S / 2;
        if (dy * dy >= 24 * 24) continue;                                         // never in range on this row
        if (px < c - 23) { if (s > 0) bound = min(bound, c - 24); }               // |x - c| <= 23 contains every x in range
        else if (px > c + 23) { if (s < 0) bound = max(bound, c + 24); }
        else return behind;
    }
    return bound;
}

// The middle of a drop / jump option: LEFT / RIGHT ticks until the player is within 4 px of the target column
// (opts:239-244 / 305-314; a jump turns round while it stands blocked, opts:309-312).  This is tick(LEFT / RIGHT)
// (impl:305-313, 331-354) for a lane that runs alone: drops and jumps are 0.04 % of the steps under random actions but
// 35-tick serial chains, the longest of a tile and the whole of a small batch's step time.  Every probe of these ticks
// (side: impl:259-281, up_clear: impl:232-238, can_fall: impl:283-288) looks at cells within two columns and a few rows
// of the player, and doors do not move during the option: an 8 x 8-cell window of the level around the player is
// gathered once into two 64-bit registers (solid = WALL or closed door, blocked = anything but OPEN), and a probe is
// two multiply-shifts and a bit test -- no shared-memory load and no cache key on the tick's dependency chain.
// Returns 1 when aligned, 0 when the tick cap was hit, -1 when the player is about to leave the window (nothing of
// that tick has been executed: the general loop continues from the same state).
template <bool TAPE, int NI>
__device__ __forceinline__ int move_until_aligned(Env<NI> &e, const LevelBlob &L, bool jump, int s, int tpx, int &n) {
    const int c0 = min(max(pad_cell(min(e.px, tpx) - 20) - 1, 0), TSTRIDE - 8);
    const int r0 = min(max(pad_cell(e.py) - 3, 0), TSTRIDE - 8);
    uint64_t ws = 0, wn = 0;
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
        const uint32_t D = door_bits(L, e.flags, r0 + rr);
        ws |= (uint64_t)(((L.row_solid[r0 + rr] | D) >> c0) & 0xFFu) << (8 * rr);
        wn |= (uint64_t)(((L.row_nonopen[r0 + rr] | D) >> c0) & 0xFFu) << (8 * rr);
    }
    auto bit = [&](uint64_t w, int x, int y) -> uint32_t {
        return (uint32_t)(w >> ((pad_cell(y) - r0) * 8 + (pad_cell(x) - c0))) & 1u;
    };
    auto can_fall_w = [&](int x, int y) { return (bit(wn, x - 10, y) | bit(wn, x + 10, y) | bit(wn, x - 10, y + 50) | bit(wn, x + 10, y + 50)) == 0u; };
    while (abs(tpx - e.px) >= 4) {
        const int x = e.px, y = e.py;
        // every cell this tick can look at (x - 20 .. x + 20 after the move, y - 4 .. y + 54 after a 4-px fall) is in the window
        // pad_cell(x - 20) >= c0 && pad_cell(x + 20) <= c0 + 7  <=>  48 c0 - 124 <= x <= 48 c0 + 219; the rows likewise
        if ((unsigned)(x - (S * c0 - 124)) > 343u || (unsigned)(y - (S * r0 - 140)) > 325u) return -1;
        const bool blocked = (bit(ws, x + 16 * s, y + 4) | bit(ws, x + 16 * s, y + 44)) != 0u;
        const bool cf = can_fall_w(x, y);
        // Blocked in mid-air (the wall of the ledge a jump is rising along, the wall a drop falls along): the policy keeps
        // asking for the same side, the side probe keeps failing -- no move, no draw, facing unchanged (impl:305-313) -- and
        // the tick is its vertical half alone (impl:331-348) with playerx fixed.  Every probe then looks at fixed window
        // columns: their 8-bit column profiles (bit r = window row r) turn the ticks into a register loop, until the side
        // probe clears, the player lands, or the window ends (the general form above takes over from the same state).
        if (blocked && cf) {
            bool near_item = false;
#pragma unroll
            for (int i = 0; i < NI; i++) near_item |= (i < L.n_items) && abs(x - (e.ix[i] + S / 2)) < 24;
            if (!near_item) {
                auto colprof = [&](uint64_t w, int xx) -> uint32_t {
                    return (uint32_t)((((w >> (pad_cell(xx) - c0)) & 0x0101010101010101ull) * 0x0102040810204080ull) >> 56);
                };
                const uint32_t pS = colprof(ws, x + 16 * s), pN = colprof(wn, x - 10) | colprof(wn, x + 10),
                               pU = colprof(wn, x - 4) | colprof(wn, x + 4);
                auto row = [&](int v) { return pad_cell(v) - r0; };
                int yy = y, tk = ticker(e.flags), ticks = 0;
                for (;;) {
                    ticks++;                                                         // a tick that starts blocked and able to fall
                    if (tk > 0) {                                                    // impl:331-334
                        if ((((pU >> row(yy - 4)) | (pU >> row(yy - 1))) & 1u) == 0u) yy -= 4;
                        tk--;
                    } else {                                                         // impl:335-346: one pixel at a time, at most 4
                        int yd = 4;
                        do {
                            yy++; yd--;
                            if ((((pN >> row(yy)) | (pN >> row(yy + 50))) & 1u) != 0u) yd = 0;
                        } while (yd > 0);
                    }
                    if (n + ticks >= TG_TICK_CAP) break;
                    if ((unsigned)(yy - (S * r0 - 140)) > 325u) break;               // the next tick may look outside the window
                    if ((((pS >> row(yy + 4)) | (pS >> row(yy + 44))) & 1u) == 0u) break;    // side probe clear
                    if ((((pN >> row(yy)) | (pN >> row(yy + 50))) & 1u) != 0u) break;        // landed
                }
                e.py = yy; e.flags = set_ticker(e.flags, tk);
                e.total_actions += ticks; n += ticks;
                if (n >= TG_TICK_CAP) return 0;
                continue;
            }
        }
        int dir = s;
        if (jump && blocked && !cf) dir = -s;
        const bool go = (dir == s) ? !blocked : (bit(ws, x + 16 * dir, y + 4) | bit(ws, x + 16 * dir, y + 44)) == 0u;
        e.total_actions++;
        int xd = 0, yd = 0;
        if (go) {
            if (!TAPE) ensure_block(e);
            xd = noisy_cached<TAPE>(e, dir < 0);
            e.flags = (e.flags & ~(1u << F_FACING)) | (dir < 0 ? 0u : (1u << F_FACING));
        }
        const int tk = ticker(e.flags);
        if (tk > 0) {                                                                // impl:331-334
            if ((bit(wn, x - 4, y - 4) | bit(wn, x + 4, y - 4) | bit(wn, x - 4, y - 1) | bit(wn, x + 4, y - 1)) == 0u) yd = -4;
            e.flags = set_ticker(e.flags, tk - 1);
        } else if (cf) {                                                             // impl:335-337
            yd = 4;
        }
        e.px += xd;                                                                  // impl:339
        if (yd > 0 && can_fall_w(e.px, y)) {                                         // impl:341-346: one pixel at a time
            do {
                e.py++; yd--;
                if (!can_fall_w(e.px, e.py)) yd = 0;
            } while (yd > 0);
        } else {
            e.py += yd;                                                              // impl:348
        }
        bool near_x = false;
#pragma unroll
        for (int i = 0; i < NI; i++) near_x |= (i < L.n_items) && abs(e.px - (e.ix[i] + S / 2)) < 24;
        if (near_x) pickups(e, L);
        n++;
        if (n >= TG_TICK_CAP) return 0;
    }
    row_cache_drop(e);
    return 1;
}

// The end of a drop / jump option once the player is aligned with its target column: NOP ticks until it cannot fall
// (opts:233-238 / 299-304), the last of them executed after the policy has seen firm ground.  playerx does not change
// any more, so every probe looks at fixed columns and two column profiles (bit r = something at x -+ 10 / x -+ 4 in padded
// row r, closed doors included) turn them into register bit tests:
//   * while a jump is still rising (impl:331-334): one tick = 4 px up if up_clear, ticker - 1 -- a loop in registers;
//   * the fall (impl:335-348, no jump in progress): one pixel at a time, at most 4 per tick, stopping at the first y
//     where can_fall fails -- the resting y comes from the profile directly and the fall takes ceil(distance / 4) ticks,
//     plus the final one.
// No draws.  Returns false (nothing done) when a key / coin is within reach of the column (pickups happen tick by tick).
template <int NI>
__device__ __forceinline__ bool fall_to_rest(Env<NI> &e, const LevelBlob &L, int &n) {
#pragma unroll
    for (int i = 0; i < NI; i++) if (i < L.n_items && abs(e.px - (e.ix[i] + S / 2)) < 24) return false;
    const int ca = pad_cell(e.px - 10), cb = pad_cell(e.px + 10), cu = pad_cell(e.px - 4), cv = pad_cell(e.px + 4);
    uint32_t blk = L.col_nonopen[ca] | L.col_nonopen[cb], upb = L.col_nonopen[cu] | L.col_nonopen[cv];
    const uint32_t closed = (e.flags >> F_DOORS) & 63u;
    for (int d = 0; d < L.n_doors; d++) {
        const int dc = L.door_cx[d] + PAD;
        if (!((closed >> d) & 1u)) continue;
        if (dc == ca || dc == cb) blk |= 1u << (L.door_cy[d] + PAD);
        if (dc == cu || dc == cv) upb |= 1u << (L.door_cy[d] + PAD);
    }
    int tk = ticker(e.flags), py = e.py, ticks = 0;
    for (;;) {
        const bool cf = (((blk >> pad_cell(py)) | (blk >> pad_cell(py + 50))) & 1u) == 0u;   // can_fall (impl:283-288)
        if (tk == 0 && cf) {
            // first y' >= y at which the row of y' or the row of y' + 50 holds something: the padded row r starts at pixel
            // 48 * (r - PAD); the border rows are WALL, so both searches find a row
            const int ra = pad_cell(py), rb = pad_cell(py + 50);
            const uint32_t ma = blk >> ra, mb = blk >> rb;
            const int ya = S * (ra + __ffs(ma) - 1 - PAD), yb = S * (rb + __ffs(mb) - 1 - PAD) - 50;
            const int ystop = max(py, min(ya, yb));
            ticks += (ystop - py + 3) / 4 + 1;
            py = ystop;
            break;
        }
        // a NOP tick with the jump still rising, or the final tick on firm ground (which may still rise: impl:331-334)
        if (tk > 0) {
            if ((((upb >> pad_cell(py - 4)) | (upb >> pad_cell(py - 1))) & 1u) == 0u) py -= 4;   // up_clear (impl:232-238)
            tk--;
        }
        ticks++;
        if (!cf) break;
    }
    if (n + ticks > TG_TICK_CAP) return false;
    e.py = py;
    e.flags = set_ticker(e.flags, tk);
    n += ticks; e.total_actions += ticks;
    row_cache_drop(e);
    return true;
}

// The option loop in its general form (opt:20-36 with the policies of opts:74-85, 146-157, 168-173, 184-189, 231-244,
// 297-314, 367-384, 426-439, 457-460): policy step, primitive tick (impl:290-359), until the policy says done -- whose
// action is still executed.  Complete for every option; run_option_to_end enters it with whatever its straight-line
// loops left.  FAST allows the closed-form fall at the end of drops and jumps; `after_tick(n)` runs after every
// primitive tick (the per-tick frame streaming of tg_step_frames, opt:33-34, runs it with FAST = false from n = 0).
template <bool TAPE, int NI, bool INTERACT_OK, bool FAST, class F>
__device__ __forceinline__ void general_option_loop(Env<NI> &e, const LevelBlob &L, int k, int s, int tpx, int &n, F after_tick) {
    bool done = false;
    do {
        int act, lad = -1;
        const bool al = abs(tpx - e.px) < 4;              // close_enough_*  (opts:69-72 ...)
        if (k <= TG_GO_RIGHT) {                           // opts:74-85 / 146-157
            done = al; act = (s < 0) ? A_LEFT : A_RIGHT;
        } else if (k == TG_UP_LADDER) {                   // opts:168-173
            done = !ladder_probe(L, e.px, e.py, true); act = done ? A_NOP : A_UP; lad = 1;
        } else if (k == TG_DOWN_LADDER) {                 // opts:184-189
            done = !ladder_probe(L, e.px, e.py, false); act = done ? A_NOP : A_DOWN; lad = 1;
        } else if (k == TG_INTERACT) {                    // opts:457-460
            done = true; act = A_INTERACT;
        } else if (k <= TG_DOWN_RIGHT) {                  // opts:231-244 / 426-439
            if (al) { if (FAST && fall_to_rest(e, L, n)) break; done = !can_fall_m(e, L); act = A_NOP; }
            else act = (s < 0) ? A_LEFT : A_RIGHT;
        } else {                                          // opts:297-314 / 367-384
            if (n == 0) act = A_JUMP;
            else if (al) { if (FAST && fall_to_rest(e, L, n)) break; done = !can_fall_m(e, L); act = A_NOP; }
            else {
                bool blocked = !side_free_m(e, L, e.px + 16 * s);
                bool grounded = !can_fall_m(e, L);
                bool rev = grounded && blocked;
                act = ((s < 0) != rev) ? A_LEFT : A_RIGHT;
            }
        }
        tick<TAPE, NI, INTERACT_OK>(e, L, act, lad);
        n++;
        after_tick(n);
        if (n >= TG_TICK_CAP && !done) { e.flags |= 1u << F_ERROR; break; }
    } while (!done);
}

// Runs option k (already known to be runnable, target column tcx from option_setup) to termination.  Returns the
// number of primitive ticks; reward = -ticks - 4*[jump option] (impl:15-16,356-359).  `valid` = this lane has an env
// and an option to run (every lane of a warp may call it).
//
// The two straight-line forms (walkers, ladders) are built from tight loops with branch-free bodies, which is what
// the time of a step is made of (measured per 32-env chunk, tools/bench_chunks.py: a loop trip that picks one of
// "four ticks / one tick / checked tick" cost a warp ~1000 cycles, 15 us per walker chunk):
//   (1) fast-forward: eight (ladders: four) unchecked ticks per trip while every one of them is sure to start inside
//       the safe stretch; the walkers compute two Philox blocks side by side;
//   (2) the rest of the stretch from the cached block, branch-free: the moves are positive, so "tick k starts inside
//       the stretch" is a comparison of the prefix sums with the distance left (consume_block);
//   (3) the tick around an event (target reached, side blocked, item in range, a probed row changes), checked.
template <bool TAPE, int NI, bool INTERACT_OK = true>
__device__ __forceinline__ int run_option_to_end(Env<NI> &e, const LevelBlob &L, int k, int tcx, bool valid = true) {
    const int tpx = tcx * S + S / 2;
    const int s = (k == TG_GO_LEFT || k == TG_DOWN_LEFT || k == TG_JUMP_LEFT) ? -1 : 1;
    int n = 0;
    bool done = false;           // the option has terminated
    // ---- go_left / go_right: straight-line form of the tick for the common situation: no jump in progress and ground
    // under the feet.  Then tick(LEFT/RIGHT) (impl:290-359) reduces to: count the action, move by noisy() if the side
    // probe is free (and face that way), pick up.  playery never changes, so the two row masks are loop constants.  Any
    // other situation (ticker > 0, can_fall) leaves the loop *before* the tick and the general loop below takes over.
    // Inside [px .. bound] (walk_safe_bound) a tick is just "count, move by noisy()": no probe can change, the target is
    // not reached, nothing can be picked up.
    // The drop options walk the same way until the ground ends or the target column is reached (opts:231-244 /
    // 426-439 return LEFT / RIGHT until aligned): they take the unchecked part of the loop and leave before (3).
    const bool drop = k == TG_DOWN_LEFT || k == TG_DOWN_RIGHT;
    if (valid && (k <= TG_GO_RIGHT || drop)) {
        if (e.m_py != fall_key(e.py)) fall_cache_fill(e, L, fall_key(e.py));
        if (e.m_py_side != side_key(e.py)) side_cache_fill(e, L, side_key(e.py));
        const uint32_t mside = e.m_side, mfall = e.m_fall;
        const bool neg = s < 0;
        int bound = ticker(e.flags) == 0 ? walk_safe_bound(e, L, s, tpx, mside, mfall) : e.px - s;
        for (;;) {
            const int n0 = n;
            if (!TAPE) {
                // (1) the largest move is 4 px: eight ticks all start inside the stretch when 28 px further on one still does
                while (s * (bound - e.px) >= 32 && ((e.draws - e.d0) & 3u) == 0u && n <= TG_TICK_CAP - 8) {
                    const uint32_t b = (e.draws - e.d0) >> 2;
                    uint32_t x0, x1, x2, x3, y0, y1, y2, y3;
                    philox4x32_10_x2(e.d0, b, e.id_lo, e.id_hi, e.key0, e.key1, x0, x1, x2, x3, y0, y1, y2, y3);
                    const int sr = (noisy_r_from_w(x0) + noisy_r_from_w(x1)) + (noisy_r_from_w(x2) + noisy_r_from_w(x3)) +
                                   (noisy_r_from_w(y0) + noisy_r_from_w(y1)) + (noisy_r_from_w(y2) + noisy_r_from_w(y3));
                    e.px += (neg ? -32 : 16) + sr;                                 // eight moves of noisy(-4) / noisy(+4)
                    e.draws += 8; n += 8;
                }
                // (2)
                for (;;) {
                    const int D = s * (bound - e.px) - 4;
                    if (D < 0 || n > TG_TICK_CAP - 4) break;
                    ensure_block(e);
                    int moved;
                    n += consume_block(e, neg, D, 4, moved);
                    e.px += neg ? -moved : moved;
                }
            } else {
                while (s * (bound - e.px) >= 4 && n < TG_TICK_CAP) { e.px += noisy<true>(e, neg); n++; }
            }
            if (n != n0) e.flags = (e.flags & ~(1u << F_FACING)) | (neg ? 0u : (1u << F_FACING));
            if (n >= TG_TICK_CAP) { e.flags |= 1u << F_ERROR; done = true; break; }
            // (3)
            if (drop) break;
            if (ticker(e.flags) != 0 ||
                (((mfall >> pad_cell(e.px - 10)) | (mfall >> pad_cell(e.px + 10))) & 1u) == 0u) break;   // jumping or falling: the general loop
            done = abs(tpx - e.px) < 4;                   // opts:80-85: the action of the final policy step still runs
            const int len0 = bag_len(e.flags);
            if (((mside >> pad_cell(e.px + 16 * s)) & 1u) == 0u) {
                if (!TAPE) ensure_block(e);
                e.px += noisy_cached<TAPE>(e, neg);
                e.flags = (e.flags & ~(1u << F_FACING)) | (neg ? 0u : (1u << F_FACING));
            }
            pickups(e, L);
            n++;
            if (done) break;
            if (n >= TG_TICK_CAP) { e.flags |= 1u << F_ERROR; done = true; break; }
            if (s * (e.px - bound) > 0 || bag_len(e.flags) != len0) bound = walk_safe_bound(e, L, s, tpx, mside, mfall);
        }
        e.total_actions += n;
    }
    // ---- up_ladder / down_ladder (opts:160-189 + impl:290-359), straight-line form.  playerx never changes, so the
    // two probes of every tick look at fixed columns: a column profile per probe (bit r = the probe's cells of padded
    // row r) turns them into register bit tests -- lad: LADDER at x -+ 12 (impl:240-257), blk: anything but OPEN at
    // x -+ 10, closed doors included (impl:283-288; doors cannot move during the option).  While the ladder probe holds
    // and the player cannot fall, tick(UP/DOWN) is: count the action, move by noisy(), pick up.  The probes only change
    // when one of the probed pixel rows (y-4, y, y+44, y+50 going up; y, y+50, y+51 going down) enters another cell row:
    // `room` is the distance to the first such change, the stretch inside which ticks run unchecked.  A falling player
    // leaves the loop *before* the tick and the general loop below takes over from the same state; the NOP tick that
    // ends the option (opts:168-173 / 184-189) is taken here.
    if (valid && (k == TG_UP_LADDER || k == TG_DOWN_LADDER) && ticker(e.flags) == 0) {
        const bool up = (k == TG_UP_LADDER);
        const int c0 = pad_cell(e.px - 12), c1 = pad_cell(e.px + 12), ca = pad_cell(e.px - 10), cb = pad_cell(e.px + 10);
        const uint32_t lad = L.col_ladder[c0] | L.col_ladder[c1];
        uint32_t blk = L.col_nonopen[ca] | L.col_nonopen[cb];
        {
            const uint32_t closed = (e.flags >> F_DOORS) & 63u;
            for (int d = 0; d < L.n_doors; d++) {
                const int dc = L.door_cx[d] + PAD;
                if (((closed >> d) & 1u) && (dc == ca || dc == cb)) blk |= 1u << (L.door_cy[d] + PAD);
            }
        }
        bool near_item = false;                     // an item in this column's pick-up range (impl:350-354)?  else no pickups at all
#pragma unroll
        for (int i = 0; i < NI; i++) near_item |= (i < L.n_items) && abs(e.px - (e.ix[i] + S / 2)) < 24;
        for (;;) {
            const int py = e.py;
            const int r1 = pad_cell(py);
            const uint32_t probe = up ? ((lad >> pad_cell(py - 4)) | (lad >> r1) | (lad >> pad_cell(py + 44)))
                                      : ((lad >> r1) | (lad >> min(r1 + 1, pad_cell(py + 51))) | (lad >> pad_cell(py + 51)));
            if ((((blk >> r1) | (blk >> pad_cell(py + 50))) & 1u) == 0u) break;       // can_fall (also when the final NOP tick starts a fall): the general loop
            if (!(probe & 1u) || (up && py <= 1)) {
                n++;                                // the NOP tick on firm ground (impl:290-359): count, pick up
                if (near_item) pickups(e, L);
                done = true;
                break;
            }
            int room = up ? min(min(mod48(py - 4), mod48(py)), min(mod48(py + 44), mod48(py + 50)))
                          : 47 - max(max(mod48(py), mod48(py + 50)), mod48(py + 51));
            {   // A longer stretch that is still safe: while the row of py itself holds a ladder cell (probe, impl:240-257) and
                // something that is not OPEN (no fall, impl:283-288) in the probed columns, neither predicate can change --
                // whatever the other probed rows enter.  `run` = the padded rows with both, contiguous from the row of py.
                const uint32_t both = lad & blk;
                if ((both >> r1) & 1u) {
                    if (up) {
                        const uint32_t gap = ~both & ((1u << r1) - 1u);                       // rows below index r1 without
                        const int r_top = gap ? 32 - __clz(gap) : 0;
                        room = max(room, py - S * (r_top - PAD));
                    } else {
                        const uint32_t gap = ~both & ~((2u << r1) - 1u);                      // rows above index r1 without
                        const int r_bot = gap ? __ffs(gap) - 2 : TSTRIDE - 1;
                        room = max(room, S * (r_bot - PAD) + S - 1 - py);
                    }
                }
            }
            if (up) room = min(room, py - 2);
            if (!TAPE && !near_item) {
                // (1) four ticks all start inside the stretch when 12 px further on one still does
                while (room >= 12 && ((e.draws - e.d0) & 3u) == 0u && n <= TG_TICK_CAP - 4) {
                    const uint4 o = philox_block(e.d0, (e.draws - e.d0) >> 2, e.id_lo, e.id_hi, e.key0, e.key1);
                    e.w0 = o.x; e.w1 = o.y; e.w2 = o.z; e.w3 = o.w;
                    const int sr = (noisy_r_from_w(o.x) + noisy_r_from_w(o.y)) + (noisy_r_from_w(o.z) + noisy_r_from_w(o.w));
                    const int mag = up ? 16 - sr : 8 + sr;                          // four moves of noisy(-4) / noisy(+4)
                    e.py += up ? -mag : mag; room -= mag;
                    e.draws += 4; n += 4;
                }
                // (2)
                while (room >= 0 && n <= TG_TICK_CAP - 4) {
                    ensure_block(e);
                    int moved;
                    n += consume_block(e, up, room, 4, moved);
                    e.py += up ? -moved : moved; room -= moved;
                }
            }
            while (room >= 0 && n < TG_TICK_CAP) {      // draw tape, an item within reach, or close to the tick cap: one tick at a time
                const int dlt = noisy<TAPE>(e, up);
                e.py += dlt; room -= abs(dlt);
                n++;
                if (near_item) pickups(e, L);
            }
            if (n >= TG_TICK_CAP) { e.flags |= 1u << F_ERROR; done = true; break; }
        }
        e.total_actions += n;
    }
    // ---- drops and jumps: the JUMP tick (opts:297-300), then LEFT / RIGHT ticks until aligned; the fall at the end is
    // taken in closed form by the loop below (fall_to_rest)
    if (valid && !done && k >= TG_DOWN_LEFT) {
        if (k >= TG_JUMP_LEFT && n == 0) { tick<TAPE, NI, false>(e, L, A_JUMP); n++; }
        if (move_until_aligned<TAPE>(e, L, k >= TG_JUMP_LEFT, s, tpx, n) == 0) { e.flags |= 1u << F_ERROR; done = true; }
    }
    // ---- everything else, and what the loops above left: policy step + general tick, one lane at its own pace ----
    if (valid && !done) general_option_loop<TAPE, NI, INTERACT_OK, true>(e, L, k, s, tpx, n, [](int) {});
    return n;
}

// Rough number of primitive ticks option k will take from this state (only used to group
// environments of similar length into the same warp; any value is correct).
template <int NI>
__device__ __forceinline__ int estimate_ticks(const Env<NI> &e, const LevelBlob &L, int k, int tcx) {
    const uint32_t f = e.flags;
    const int walk = (abs(tcx * S + S / 2 - e.px) + 2) / 3;          // noisy() moves 3 px per tick on average
    int pcx, pcy; player_cell(e, pcx, pcy);
    switch (k) {
    case TG_GO_LEFT: case TG_GO_RIGHT: return walk + 1;
    case TG_UP_LADDER: {
        int yc = pcy, n = 0;
        while (n < TSTRIDE && (type_c(L, f, pcx, yc) == T_LADDER || type_c(L, f, pcx, yc - 1) == T_LADDER)) { yc--; n++; }
        return max(2, (e.py - (yc + 1) * S + S) / 3);
    }
    case TG_DOWN_LADDER: {
        int yc = pcy + 1, n = 0;
        while (n < TSTRIDE && type_c(L, f, pcx, yc) == T_LADDER) { yc++; n++; }
        return max(2, (yc * S - e.py - S) / 3 + 2);
    }
    case TG_INTERACT: return 1;
    case TG_DOWN_LEFT: case TG_DOWN_RIGHT: {
        int yc = pcy + 1, n = 0;
        while (n < TSTRIDE && type_c(L, f, tcx, yc) == T_OPEN) { yc++; n++; }
        return walk + 12 * n + 2;
    }
    default: return 24 + walk;                                       // jump: 22-23 ticks of lift + drift
    }
}

template <int NI>
__device__ __forceinline__ bool is_done(const Env<NI> &e, const LevelBlob &L) {    // tg:95
    return has_gold(L, e.flags) && floordiv48(e.py + S / 2) == 0;
}

// The plan word of an env (tg_types.h PL_*): can_run of the nine options (tg:83-89), the done predicate (tg:95),
// "the reference would raise" for the two drop options, and a length class per option for the step kernel's sort.
// A function of (position, flags, item positions) only; every kernel that stores an env's state stores its plan.
// All nine can_run predicates are evaluated together on the row masks of the player's cell row and its two
// neighbours (the same formulation as walk_setup / ladder_probe; option_setup keeps the per-option form the
// option lanes use, and the parity tests compare both with the oracle's masks).  Out of line, scalar arguments.
__device__ __forceinline__ uint32_t plan_len(int ticks) { return (uint32_t)min(ticks / 10, 11); }
template <int NI>
__device__ __noinline__ uint64_t compute_plan(const LevelBlob *Lp, uint32_t pxy, uint32_t flags, uint32_t it0, uint32_t it1,
                                              uint32_t it2, uint32_t it3) {
    const LevelBlob &L = *Lp;
    const int px = (int)(pxy & 0xFFFu), py = (int)(int16_t)(pxy >> 16);
    const int cxp = pad_cell(px), r = pad_cell(py + S / 2);                       // padded player cell (impl:441-445)
    const uint32_t closed = (flags >> F_DOORS) & 63u;
    auto doors = [&](int row) -> uint32_t { const int li = L.row_lut[row]; return li ? L.door_lut[li - 1][closed] : 0u; };
    const uint32_t D0 = doors(r), D1 = doors(r + 1);
    const uint32_t open_m = ~(L.row_nonopen[r - 1] | doors(r - 1)), open_0 = ~(L.row_nonopen[r] | D0), open_1 = ~(L.row_nonopen[r + 1] | D1);
    const uint32_t wall_0 = L.row_solid[r];                                        // WALL proper (a closed door reads DOOR)
    uint32_t lo = 0, len_lo = 0, len_hi = 0;                                      // len: 9 x 4 bits = 32 + 4
    const uint32_t items[4] = {it0, it1, it2, it3};

    // go_left / go_right (opts:23-67, :95-139), as walk_setup
    {
        uint32_t obj = L.row_static_obj[r] | D0;                                   // impl:402-409
#pragma unroll
        for (int i = 0; i < NI; i++) {
            const int icx = (int)(int16_t)(items[i] & 0xFFFFu) / S + PAD, icy = (int)(int16_t)(items[i] >> 16) / S + PAD;
            if (i < L.n_items && icy == r && icx >= 0 && icx < TSTRIDE) obj |= 1u << icx;
        }
        const uint32_t nb = wall_0 | D0 | open_1;
        const uint32_t tm = L.row_ladder[r - 1] | L.row_ladder[r + 1] | obj;
        const uint32_t okm = open_0 & ~open_1;                                     // opts:34-39
        const uint32_t candL = (tm | (nb << 1)) & ((1u << cxp) - 1u) & ~((1u << PAD) - 1u);
        if (candL) {
            const int txp = 31 - __clz(candL);
            const uint32_t range = ((2u << cxp) - 1u) & ~((1u << txp) - 1u);
            if ((okm & range) == range) { lo |= 1u << TG_GO_LEFT; len_lo |= plan_len((abs((txp - PAD) * S + S / 2 - px) + 2) / 3 + 1) << (4 * TG_GO_LEFT); }
        }
        const uint32_t candR = (tm | (nb >> 1)) & ~((2u << cxp) - 1u);
        if (candR) {
            const int txp = __ffs(candR) - 1;
            const uint32_t range = ((2u << txp) - 1u) & ~((1u << cxp) - 1u);
            if ((okm & range) == range) { lo |= 1u << TG_GO_RIGHT; len_lo |= plan_len((abs((txp - PAD) * S + S / 2 - px) + 2) / 3 + 1) << (4 * TG_GO_RIGHT); }
        }
    }
    // up_ladder / down_ladder (opts:165-166, :181-182 = impl:240-257)
    if (ladder_probe(L, px, py, true)) {
        int rr = r;
        while (rr > 0 && (((L.row_ladder[rr] | L.row_ladder[rr - 1]) >> cxp) & 1u)) rr--;
        lo |= 1u << TG_UP_LADDER; len_lo |= plan_len(max(2, (py - (rr - PAD) * S) / 3)) << (4 * TG_UP_LADDER);
    }
    if (ladder_probe(L, px, py, false)) {
        int rr = r + 1;
        while (rr < TSTRIDE - 1 && ((L.row_ladder[rr] >> cxp) & 1u)) rr++;
        lo |= 1u << TG_DOWN_LADDER; len_lo |= plan_len(max(2, ((rr - PAD) * S - py - S) / 3 + 2)) << (4 * TG_DOWN_LADDER);
    }
    // interact (opts:446-455): one tick
    {
        bool ok = false;
        for (int i = 0; i < L.n_handles; i++) ok |= near_px(px, py, L.handle_cx[i] * S, L.handle_cy[i] * S, 36 * 36);
        if (has_key(L, flags))
            for (int i = 0; i < L.n_bolts; i++) ok |= near_px(px, py, L.bolt_cx[i] * S, L.bolt_cy[i] * S, 24 * 24);
        if (ok) lo |= 1u << TG_INTERACT;
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int s = side ? 1 : -1, c1 = cxp + s, c2 = cxp + 2 * s;
        // down_left / down_right (opts:199-221, :394-416): the side cell and the one below it are OPEN; the target row is
        // the first cell that is not, and the reference raises when there is none
        if ((open_0 >> c1) & (open_1 >> c1) & 1u) {
            int rr = r + 1;
            bool err = false;
            for (;;) {
                if (((L.row_nonopen[rr] | doors(rr)) >> c1) & 1u) break;
                rr++;
                if (rr - PAD >= L.ch) { err = true; break; }
            }
            const int k = side ? TG_DOWN_RIGHT : TG_DOWN_LEFT;
            if (err) lo |= 1u << (side ? PL_ERR_DR : PL_ERR_DL);
            else { lo |= 1u << k; len_lo |= plan_len((abs((c1 - PAD) * S + S / 2 - px) + 2) / 3 + 12 * (rr - r - 1) + 2) << (4 * k); }
        }
        // jump_left / jump_right (opts:254-287, :324-357): OPEN above and diagonally above, and a landing (OPEN over WALL)
        // one or two columns away in the row above
        if ((open_m >> cxp) & (open_m >> c1) & 1u) {
            int tc = 0;
            if ((wall_0 >> c1) & 1u) tc = c1;
            else if ((open_m >> c2) & (wall_0 >> c2) & 1u) tc = c2;
            if (tc) {
                const uint32_t lc = plan_len(24 + (abs((tc - PAD) * S + S / 2 - px) + 2) / 3);
                if (side) { lo |= 1u << TG_JUMP_RIGHT; len_hi |= lc; } else { lo |= 1u << TG_JUMP_LEFT; len_lo |= lc << (4 * TG_JUMP_LEFT); }
            }
        }
    }
    if (has_gold(L, flags) && r == PAD) lo |= 1u << PL_TERM;                       // tg:95
    return (uint64_t)lo | ((uint64_t)len_lo << PL_LEN) | ((uint64_t)len_hi << (PL_LEN + 32));
}
template <int NI>
__device__ __forceinline__ uint64_t plan_of(const Env<NI> &e, const LevelBlob &L) {
    return compute_plan<NI>(&L, pack_player(e.px, e.py, 0u), e.flags, pack_xy(e.ix[0], e.iy[0]),
                            NI > 1 ? pack_xy(e.ix[NI > 1 ? 1 : 0], e.iy[NI > 1 ? 1 : 0]) : 0u,
                            NI > 2 ? pack_xy(e.ix[NI > 2 ? 2 : 0], e.iy[NI > 2 ? 2 : 0]) : 0u,
                            NI > 3 ? pack_xy(e.ix[NI > 3 ? 3 : 0], e.iy[NI > 3 ? 3 : 0]) : 0u);
}
__device__ __forceinline__ int plan_len_class(uint64_t p, int k) { return (int)((p >> (PL_LEN + 4 * k)) & 15u); }

// ---------------------------------------------------------------------------
// reset  (impl:55-73, impl:168-178, objs:106-115) -- draws: one per handle (file order), then a gauss pair
// ---------------------------------------------------------------------------
template <bool TAPE, int NI>
__device__ __forceinline__ void reset_env(Env<NI> &e, const LevelBlob &L) {
    e.flags = L.init_flags | (e.flags & (1u << F_ERROR));      // the error flag is sticky until tg_reset
    e.sticky = 0;                                              // fresh objects (impl:57-60)
    for (int h = 0; h < L.n_handles; h++) {
        bool up = (L.init_flags >> (F_HANDLES + h)) & 1u;
        e.angles[(int64_t)h * e.n] = handle_angle(up, draw<TAPE>(e));
    }
#pragma unroll
    for (int i = 0; i < NI; i++)
        if (i < L.n_items) { e.ix[i] = L.item_cx[i] * S; e.iy[i] = L.item_cy[i] * S; }
    // CPython random.gauss(mu, sigma): x2pi = random()*2pi; g2rad = sqrt(-2 log(1 - random()));
    // z = cos(x2pi)*g2rad; gauss_next = sin(x2pi)*g2rad  (the second gauss() call consumes gauss_next)
    double x2pi = __dmul_rn(draw<TAPE>(e), 6.283185307179586);
    double g2rad = __dsqrt_rn(__dmul_rn(-2.0, log(__dsub_rn(1.0, draw<TAPE>(e)))));
    double z0 = __dmul_rn(cos(x2pi), g2rad), z1 = __dmul_rn(sin(x2pi), g2rad);
    int nx = (int)__dmul_rn(z0, 2.0);                            // int(gauss(0, 48/24))   impl:170
    int ny = (int)fabs(__dmul_rn(z1, 48.0 / 36.0));              // int(abs(gauss(0, 48/36)))  impl:171
    e.px = L.start_px + nx; e.py = L.start_py + ny;              // impl:176
    if (L.start_px < 0) { e.px = 0; e.py = 0; }                  // impl:178 (level without a free cell)
    e.total_actions = 0;
}

// ---------------------------------------------------------------------------
// observation  (impl:368-378 + objs get_state): float64 arithmetic, stored as float32
// ---------------------------------------------------------------------------
// One table look-up per coordinate: lut = this level's pair of tables (BatchView::obs_lut), lut[0][v + S] =
// float32(v / W), lut[1][v + S] = float32(v / H) -- the reference's float64 quotient (impl:370-371, objs get_state)
// rounded to float32, filled on the host.  Coordinates outside the tables (only reachable through injected item
// positions) take one IEEE float division, which gives the same value: for these small integers the correctly
// rounded float32 quotient equals the float64 quotient rounded to float32 (the quotient is a multiple of
// 1/(W*2^k) away from every rounding boundary, far more than 2^-53).  The slot layout comes from the level's
// observation program (LevelBlob::obs_prog), so there is no per-object dispatch here.
static __device__ __noinline__ float obs_quot_far(int v, int extent) { return __fdiv_rn((float)v, (float)extent); }   // out of line: the division is ~40 instructions and write_obs has a dozen sites
__device__ __forceinline__ float obs_quot(const float *__restrict__ lut, int v, int extent) {
    const unsigned idx = (unsigned)(v + S);
    return idx < (unsigned)OBS_LUT_N ? __ldg(lut + idx) : obs_quot_far(v, extent);
}
// Writes the row to `o` and, when given, to `o2` (the sparse record of tg_step_host_sparse): every value is computed once.
template <int NI>
__device__ __forceinline__ void write_obs(const Env<NI> &e, const LevelBlob &L, const float *__restrict__ lut,
                                          float *__restrict__ o, int obs_dim, float *__restrict__ o2 = nullptr, bool full = true) {
    const int W = L.cw * S, H = L.ch * S;
    const int nk = L.obs_dim;
    auto put = [&](int k, float v) { if (o) o[k] = v; if (o2) o2[k] = v; };
    if (!full) {                     // only the player moved (the row holds the env's previous observation): slots 0, 1 (impl:370-371)
        put(0, obs_quot(lut, e.px, W)); put(1, obs_quot(lut + OBS_LUT_N, e.py, H));
        return;
    }
    if (L.obs_prog[31] == 1) {       // the shipped layout's slot order (two handles, key, bolt, gold): straight-line code
        const float *lx = lut, *ly = lut + OBS_LUT_N;
        const double a0 = e.angles[0], a1 = e.angles[e.n];
        put(0, obs_quot(lx, e.px, W)); put(1, obs_quot(ly, e.py, H));
        put(2, (float)a0); put(3, (float)a1);
        put(4, obs_quot(lx, e.ix[0], W)); put(5, obs_quot(ly, e.iy[0], H));
        put(6, ((e.flags >> F_BOLTS) & 1u) ? 1.0f : 0.0f);
        put(7, obs_quot(lx, e.ix[NI > 1 ? 1 : 0], W)); put(8, obs_quot(ly, e.iy[NI > 1 ? 1 : 0], H));
        for (int k = 9; k < obs_dim; k++) put(k, 0.0f);
        return;
    }
#pragma unroll 1
    for (int k = 0; k < nk; k++) {
        const int p = L.obs_prog[k], op = p >> 4, i = p & 15;
        float v;
        if (op >= OP_IX) {
            int x = 0, y = 0;
#pragma unroll
            for (int q = 0; q < NI; q++) if (q == i) { x = e.ix[q]; y = e.iy[q]; }
            v = (op == OP_IX) ? obs_quot(lut, x, W) : obs_quot(lut + OBS_LUT_N, y, H);
        } else if (op == OP_ANGLE) v = (float)e.angles[(int64_t)i * e.n];
        else if (op == OP_BOLT) v = ((e.flags >> (F_BOLTS + i)) & 1u) ? 1.0f : 0.0f;
        else v = (op == OP_PX) ? obs_quot(lut, e.px, W) : obs_quot(lut + OBS_LUT_N, e.py, H);
        put(k, v);
    }
    for (int k = nk; k < obs_dim; k++) put(k, 0.0f);
}

// ---------------------------------------------------------------------------
// init_with_state (impl:447-481) with its quirks: -99 keeps the current value through a float64 round trip
// (int(float(px)/W*W) can lose a pixel), every key / gold / bolt reads the FIRST slot carrying its name
// (desc.index), facing is forced right, bag / ticker / total_actions are untouched, handle.set_angle
// propagates triggers (objs:133-143; targets redraw their angle) and leaves the handle flagged.
// ---------------------------------------------------------------------------
template <bool TAPE, int NI>
__device__ void init_with_state_env(Env<NI> &e, const LevelBlob &L, const double *__restrict__ in) {
    const double W = (double)(L.cw * S), H = (double)(L.ch * S);
    double st[2 + 2 * TG_MAX_OBJECTS];
    int first_key = -1, first_gold = -1, first_bolt = -1;
    st[0] = (double)e.px / W; st[1] = (double)e.py / H;                   // impl:368-378 (current vector)
    for (int o = 0; o < L.n_objs; o++) {
        const int kind = L.obj_kind[o], i = L.obj_idx[o], k = L.obj_obs[o];
        if (kind == TG_HANDLE) st[k] = e.angles[(int64_t)i * e.n];
        else if (kind == TG_BOLT) { st[k] = ((e.flags >> (F_BOLTS + i)) & 1u) ? 1.0 : 0.0; if (first_bolt < 0) first_bolt = k; }
        else if (kind == TG_KEY || kind == TG_GOLD) {
            int x = 0, y = 0;
#pragma unroll
            for (int q = 0; q < NI; q++) if (q == i) { x = e.ix[q]; y = e.iy[q]; }
            st[k] = (double)x / W; st[k + 1] = (double)y / H;
            if (kind == TG_KEY) { if (first_key < 0) first_key = k; } else if (first_gold < 0) first_gold = k;
        }
    }
    for (int v = 0; v < L.obs_dim; v++) { const double g = in[v]; if (g != -99.0) st[v] = g; }
    e.flags |= 1u << F_FACING;
    e.px = (int)__dmul_rn(st[0], W); e.py = (int)__dmul_rn(st[1], H);
    e.px = min(max(e.px, 0), L.cw * S - 1); e.py = min(max(e.py, -(S - 1)), L.ch * S - 1);   // probe invariants
    for (int o = 0; o < L.n_objs; o++) {
        const int kind = L.obj_kind[o], i = L.obj_idx[o], k = L.obj_obs[o];
        if (kind == TG_KEY || kind == TG_GOLD) {                             // objs:40-44 move_to_xy
            const int slot = (kind == TG_KEY) ? first_key : first_gold;
            const int x = (int)__dmul_rn(st[slot], W), y = (int)__dmul_rn(st[slot + 1], H);
#pragma unroll
            for (int q = 0; q < NI; q++) if (q == i) { e.ix[q] = x; e.iy[q] = y; }
        } else if (kind == TG_HANDLE) {                                      // objs:133-143 set_angle
            const double ang = st[k];
            const bool old = (e.flags >> (F_HANDLES + i)) & 1u, up = !(ang <= 0.15);
            e.angles[(int64_t)i * e.n] = ang;
            e.flags = (e.flags & ~(1u << (F_HANDLES + i))) | ((up ? 1u : 0u) << (F_HANDLES + i));
            if (up != old) trigger_dfs<TAPE>(e, L, o, up);
            e.sticky |= 1u << i;                                             // impl:473
        } else if (kind == TG_BOLT) {
            set_val<TAPE>(e, L, o, st[first_bolt] > 0.5);
        }
    }
    row_cache_drop(e);
}

// ---------------------------------------------------------------------------
// packed state <-> registers
// ---------------------------------------------------------------------------
// position / flags / items only (enough for can_run and the option targets)
template <int NI>
__device__ __forceinline__ void load_core(Env<NI> &e, const BatchView &B, int64_t i, uint4 c) {
    e.px = core_px(c.x); e.py = hi16(c.x); e.sticky = core_sticky(c.x); e.flags = c.y;
    e.ix[0] = lo16(c.z); e.iy[0] = hi16(c.z);
    if (NI > 1) { e.ix[1] = lo16(c.w); e.iy[1] = hi16(c.w); }
    if (NI > 2) {
        uint2 h = B.items23[i];
        e.ix[2] = lo16(h.x); e.iy[2] = hi16(h.x);
        if (NI > 3) { e.ix[3] = lo16(h.y); e.iy[3] = hi16(h.y); }
    }
}

template <int NI>
__device__ __forceinline__ void load_core(Env<NI> &e, const BatchView &B, int64_t i) { load_core(e, B, i, B.core[i]); }

template <int NI>
__device__ __forceinline__ void load_env(Env<NI> &e, const BatchView &B, int64_t i, uint4 &acct) {
    uint4 c = B.core[i];
    acct = B.acct[i];
    e.px = core_px(c.x); e.py = hi16(c.x); e.sticky = core_sticky(c.x); e.flags = c.y;
    e.ix[0] = lo16(c.z); e.iy[0] = hi16(c.z);
    if (NI > 1) { e.ix[1] = lo16(c.w); e.iy[1] = hi16(c.w); }
    if (NI > 2) {
        uint2 h = B.items23[i];
        e.ix[2] = lo16(h.x); e.iy[2] = hi16(h.x);
        if (NI > 3) { e.ix[3] = lo16(h.y); e.iy[3] = hi16(h.y); }
    }
    e.draws = acct.x; e.total_actions = acct.w;
    row_cache_drop(e); e.m_fall = 0; e.m_side = 0;
    e.d0 = acct.x; e.w0 = e.w1 = e.w2 = e.w3 = 0;
    e.key0 = B.seed_lo; e.key1 = B.seed_hi;
    uint64_t id = (uint64_t)(B.first_env_id + i);
    e.id_lo = (uint32_t)id; e.id_hi = (uint32_t)(id >> 32);
    e.tape = B.tape ? B.tape + B.tape_off[i] : nullptr;
    e.tape_len = B.tape ? (uint32_t)(B.tape_off[i + 1] - B.tape_off[i]) : 0u;
    e.angles = B.angles + i; e.n = B.n;
}

template <int NI>
__device__ __forceinline__ void store_core(const Env<NI> &e, const BatchView &B, int64_t i) {
    uint4 c;
    c.x = pack_player(e.px, e.py, e.sticky); c.y = e.flags;
    c.z = pack_xy(e.ix[0], e.iy[0]);
    c.w = (NI > 1) ? pack_xy(e.ix[1], e.iy[1]) : 0u;
    B.core[i] = c;
    if (NI > 2) {
        uint2 h;
        h.x = pack_xy(e.ix[2], e.iy[2]);
        h.y = (NI > 3) ? pack_xy(e.ix[3], e.iy[3]) : 0u;
        B.items23[i] = h;
    }
}
template <int NI>
__device__ __forceinline__ void store_acct(const Env<NI> &e, const BatchView &B, int64_t i, uint4 acct) {
    acct.x = e.draws; acct.w = e.total_actions;
    B.acct[i] = acct;
}

template <int NI>
__device__ __forceinline__ void store_env(const Env<NI> &e, const BatchView &B, int64_t i, uint4 acct) {
    uint4 c;
    c.x = pack_player(e.px, e.py, e.sticky); c.y = e.flags;
    c.z = pack_xy(e.ix[0], e.iy[0]);
    c.w = (NI > 1) ? pack_xy(e.ix[1], e.iy[1]) : 0u;
    B.core[i] = c;
    if (NI > 2) {
        uint2 h;
        h.x = pack_xy(e.ix[2], e.iy[2]);
        h.y = (NI > 3) ? pack_xy(e.ix[3], e.iy[3]) : 0u;
        B.items23[i] = h;
    }
    acct.x = e.draws; acct.w = e.total_actions;
    B.acct[i] = acct;
}

}  // namespace tg
