// TEST INFRASTRUCTURE (CPU suite only; never built into or loaded by the product): the device functions of
// gym_treasure_game_b200/csrc/tg_device.cuh -- option execution, primitive tick, plan word, reset, observation row --
// compiled for the HOST by g++ behind a handful of intrinsic shims, and driven one env at a time the way
// tg_step_kernel / tg_reset_kernel drive them (csrc/tg_step.cu: phase A's idle path, phase B's option lanes, phase C's
// reset pass).  tests/test_device_code_on_host.py compares the result step by step with the C oracle, so the kernel's
// arithmetic is under differential test in a container without a GPU.  What this does NOT cover is everything
// that is parallel in the kernel (tiles, the in-tile sort, the chunk queue, records, statistics): the -m gpu tests do.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>        // vector types and the (empty, for a host compiler) __device__ / __forceinline__ qualifiers

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
using std::max;
using std::min;
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline double __dadd_rn(double a, double b) { return a + b; }      // compiled with -ffp-contract=off: no FMA
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline unsigned long long __double2ull_rz(double a) { return a <= 0.0 ? 0ull : (unsigned long long)a; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
template <class T> static inline T __ldg(const T *p) { return *p; }

#include "../../gym_treasure_game_b200/csrc/tg_device.cuh"

using namespace tg;

namespace {

struct Batch {
    int ni = 2;
    int64_t n = 0;
    LevelBlob level;
    std::vector<uint32_t> closure;
    std::vector<uint4> core, acct;
    std::vector<uint64_t> plan;
    std::vector<uint32_t> ep_start;
    std::vector<uint2> items23;
    std::vector<double> angles;
    std::vector<float> lut;
    std::vector<double> tape;          // parity mode (tg_set_draw_tape): recorded reference draws, per env
    std::vector<int64_t> tape_off;
    uint32_t counter = 0;              // BatchView::step_counter[0]
    bool obs_stale = false;            // a call changed state without writing rows (the library's obs_sync == NULL): the next
                                       // step rewrites every row, as tg_obs_kernel does after the step kernel
    BatchView B{};
};

template <bool TAPE, int NI>
void reset_one(Batch &b, int64_t i, float *obs) {                                   // tg_reset_kernel (mask == NULL)
    const LevelBlob &L = b.level;
    Env<NI> e; uint4 acct;
    load_env(e, b.B, i, acct);
    e.flags &= ~(1u << F_ERROR);
    reset_env<TAPE>(e, L);
    acct.y = 0;
    store_env(e, b.B, i, acct);
    b.plan[i] = plan_of(e, L);
    b.ep_start[i] = b.counter;
    if (obs) write_obs(e, L, b.lut.data(), obs + i * b.B.obs_dim, b.B.obs_dim);
}

// phase C for one env: `drawn` uniforms were consumed by the option that ended the episode in this call (0: idle env)
template <bool TAPE, int NI>
void reset_in_step(Batch &b, int64_t i, uint32_t drawn, uint32_t t_now, float *obs) {
    const LevelBlob &L = b.level;
    Env<NI> e; uint4 acct;
    load_env(e, b.B, i, acct);
    acct.y = 0; acct.z = 0;
    e.d0 = e.draws - drawn;
    if (!TAPE && (drawn & 3u)) { const uint4 o = philox_block(e.d0, drawn >> 2, e.id_lo, e.id_hi, e.key0, e.key1); e.w0 = o.x; e.w1 = o.y; e.w2 = o.z; e.w3 = o.w; }
    reset_env<TAPE>(e, L);
    store_env(e, b.B, i, acct);
    b.plan[i] = plan_of(e, L);
    b.ep_start[i] = t_now + 1u;
    if (obs) write_obs(e, L, b.lut.data(), obs + i * b.B.obs_dim, b.B.obs_dim);
}

template <bool TAPE, int NI>
void step_one(Batch &b, int64_t i, int action, uint32_t t_now, float *obs, float *reward, uint8_t *done, uint8_t *ran, int32_t *ticks) {
    const LevelBlob &L = b.level;
    const BatchView &B = b.B;
    const uint32_t max_steps = B.max_steps > 0 ? (uint32_t)B.max_steps : 0xFFFFFFFFu;
    const int od = B.obs_dim;
    const uint32_t a = (uint32_t)action < (uint32_t)TG_NUM_OPTIONS ? (uint32_t)action : 12u;
    const uint64_t plan0 = b.plan[i];
    const uint32_t lo = (uint32_t)plan0;
    const uint32_t steps = t_now + 1u - b.ep_start[i];
    if (!((lo >> a) & 1u)) {                                                        // phase A: the option cannot run (opt:22-23)
        const uint32_t d = ((lo >> PL_TERM) & 1u) | (steps >= max_steps ? (uint32_t)TG_DONE_TRUNCATED : 0u);
        if ((lo >> (a + (uint32_t)(PL_ERR_DL - TG_DOWN_LEFT))) & (((3u << TG_DOWN_LEFT) >> a) & 1u))
            b.core[i].y |= 1u << F_ERROR;                                            // the reference would raise
        reward[i] = 0.f; done[i] = (uint8_t)d; ran[i] = 0; ticks[i] = 0;
        if (d && B.auto_reset) reset_in_step<TAPE, NI>(b, i, 0u, t_now, obs);
        return;
    }
    Env<NI> e; uint4 acct;
    uint32_t drawn_i = 0, flags0 = 0;
    int tcx = 0;
    if (a == TG_INTERACT) {                                                         // interact_option_mem
        load_env(e, B, i, acct);
        tick<TAPE, NI, true>(e, L, A_INTERACT);
        store_env(e, B, i, acct);
        drawn_i = e.draws - e.d0;
    }
    load_env(e, B, i, acct);
    if (a == TG_INTERACT) e.d0 = e.draws - drawn_i;
    else { bool err; option_setup(e, L, (int)a, tcx, err); flags0 = e.flags; }
    int n = run_option_to_end<TAPE, NI, false>(e, L, (int)a, tcx, a != TG_INTERACT);
    if (a == TG_INTERACT) n = 1;
    const bool jump = a >= TG_JUMP_LEFT;
    const int r = -n - (jump ? 4 : 0);
    acct.y = (uint32_t)((int)acct.y + r);
    const bool term = is_done(e, L);
    const uint32_t d = (term ? TG_DONE_TERMINATED : 0) | (steps >= max_steps ? TG_DONE_TRUNCATED : 0);
    if (d && B.auto_reset) {
        acct.y = 0;
        store_env(e, B, i, acct);
        reset_in_step<TAPE, NI>(b, i, min(e.draws - e.d0, 0xFFFFu), t_now, obs);
    } else {
        store_env(e, B, i, acct);
        b.plan[i] = plan_of(e, L);
        if (obs) write_obs(e, L, b.lut.data(), obs + i * od, od, nullptr,
                           a == TG_INTERACT || (((e.flags ^ flags0) >> F_ERROR) | ((e.flags >> F_ERROR) & 1u)) != 0u);
    }
    reward[i] = (float)r; done[i] = (uint8_t)d; ran[i] = 1; ticks[i] = n;
}

template <int NI>
void obs_one(Batch &b, int64_t i, float *obs) {                                     // tg_obs_kernel
    Env<NI> e; uint4 acct;
    load_env(e, b.B, i, acct);
    write_obs(e, b.level, b.lut.data(), obs + i * b.B.obs_dim, b.B.obs_dim);
}

template <bool TAPE, int NI>
void init_with_state_one(Batch &b, int64_t i, const double *states) {               // tg_init_with_state_kernel
    Env<NI> e; uint4 acct;
    load_env(e, b.B, i, acct);
    init_with_state_env<TAPE>(e, b.level, states + i * b.B.obs_dim);
    store_env(e, b.B, i, acct);
    b.plan[i] = plan_of(e, b.level);
}

template <bool TAPE, int NI>
void primitive_one(Batch &b, int64_t i, int a, uint32_t t_now, float *obs, float *reward, uint8_t *done) {   // tg_primitive_kernel
    const LevelBlob &L = b.level;
    const BatchView &B = b.B;
    Env<NI> e; uint4 acct;
    load_env(e, B, i, acct);
    tick<TAPE>(e, L, a);
    const int r = (a == A_JUMP) ? -5 : -1;
    acct.y = (uint32_t)((int)acct.y + r);
    const uint32_t steps = t_now + 1u - b.ep_start[i];
    const bool term = is_done(e, L);
    const bool trunc = B.max_steps > 0 && steps >= (uint32_t)B.max_steps;
    const int d = (term ? TG_DONE_TERMINATED : 0) | (trunc ? TG_DONE_TRUNCATED : 0);
    if (d && B.auto_reset) { reset_env<TAPE>(e, L); acct.y = 0; b.ep_start[i] = t_now + 1u; }
    store_env(e, B, i, acct);
    b.plan[i] = plan_of(e, L);
    if (obs) write_obs(e, L, b.lut.data(), obs + i * B.obs_dim, B.obs_dim);
    reward[i] = (float)r; done[i] = (uint8_t)d;
}

template <class F2, class F4>
void for_each_env(Batch &b, F2 f2, F4 f4) { for (int64_t i = 0; i < b.n; i++) { if (b.ni > 2) f4(i); else f2(i); } }

}  // namespace

extern "C" {

size_t hostdev_blob_size() { return sizeof(LevelBlob); }

// blob: sizeof(LevelBlob) bytes of a compiled level (the head of the product's tg_level); closure: 0 = walk the trigger
// graph, else [n_objs * 2 << 13] entries (tg_debug_level_closure).  frame_w / frame_h: the level's frame size in pixels.
void *hostdev_create(const void *blob, const uint32_t *closure, int64_t n_closure, int64_t n, int64_t first_env_id,
                     uint64_t seed, int max_steps, int auto_reset, int frame_w, int frame_h) {
    Batch *b = new Batch();
    std::memcpy(&b->level, blob, sizeof(LevelBlob));
    if (closure && n_closure > 0) { b->closure.assign(closure, closure + n_closure); b->level.closure = b->closure.data(); }
    else b->level.closure = nullptr;
    b->n = n; b->ni = b->level.n_items > 2 ? 4 : 2;
    b->core.assign(n, make_uint4(0, 0, 0, 0)); b->acct.assign(n, make_uint4(0, 0, 0, 0));
    b->plan.assign(n, 0); b->ep_start.assign(n, 0); b->items23.assign(n, make_uint2(0, 0));
    b->angles.assign((size_t)TG_MAX_HANDLES * n, 0.0);
    b->lut.resize((size_t)2 * OBS_LUT_N);
    for (int v = 0; v < OBS_LUT_N; v++) {                                            // tg_create's quotient tables
        b->lut[v] = (float)((double)(v - S) / (double)frame_w);
        b->lut[OBS_LUT_N + v] = (float)((double)(v - S) / (double)frame_h);
    }
    BatchView &B = b->B;
    B.n = n; B.first_env_id = first_env_id; B.r_begin = 0; B.r_count = n;
    B.core = b->core.data(); B.acct = b->acct.data(); B.plan = b->plan.data(); B.ep_start = b->ep_start.data();
    B.items23 = b->ni > 2 ? b->items23.data() : nullptr; B.angles = b->angles.data();
    B.level_id = nullptr; B.levels = &b->level; B.n_levels = 1; B.obs_dim = b->level.obs_dim;
    B.max_steps = max_steps; B.auto_reset = auto_reset;
    B.seed_lo = (uint32_t)seed; B.seed_hi = (uint32_t)(seed >> 32);
    B.obs_lut = b->lut.data();
    return b;
}

void hostdev_destroy(void *h) { delete static_cast<Batch *>(h); }

// parity mode, like tg_set_draw_tape: env i draws tape[off[i] .. off[i + 1]) instead of Philox words
void hostdev_set_tape(void *h, const double *tape, const int64_t *off) {
    Batch &b = *static_cast<Batch *>(h);
    b.tape_off.assign(off, off + b.n + 1);
    b.tape.assign(tape, tape + off[b.n]);
    if (b.tape.empty()) b.tape.push_back(0.0);
    b.B.tape = b.tape.data(); b.B.tape_off = b.tape_off.data();
}

void hostdev_reset(void *h, float *obs) {
    Batch &b = *static_cast<Batch *>(h);
    const bool tape = b.B.tape != nullptr;
    for (int64_t i = 0; i < b.n; i++) {
        if (tape) { if (b.ni > 2) reset_one<true, 4>(b, i, obs); else reset_one<true, 2>(b, i, obs); }
        else { if (b.ni > 2) reset_one<false, 4>(b, i, obs); else reset_one<false, 2>(b, i, obs); }
    }
}

// one TreasureGame.step for every env; obs is the caller's persistent [n][obs_dim] buffer (rows are updated in place,
// like the bound buffer of tg_step)
void hostdev_step(void *h, const int32_t *actions, float *obs, float *reward, uint8_t *done, uint8_t *ran, int32_t *ticks) {
    Batch &b = *static_cast<Batch *>(h);
    const uint32_t t_now = b.counter;
    const bool tape = b.B.tape != nullptr;
    for (int64_t i = 0; i < b.n; i++) {
        if (tape) {
            if (b.ni > 2) step_one<true, 4>(b, i, actions[i], t_now, obs, reward, done, ran, ticks);
            else step_one<true, 2>(b, i, actions[i], t_now, obs, reward, done, ran, ticks);
        } else {
            if (b.ni > 2) step_one<false, 4>(b, i, actions[i], t_now, obs, reward, done, ran, ticks);
            else step_one<false, 2>(b, i, actions[i], t_now, obs, reward, done, ran, ticks);
        }
    }
    b.counter = t_now + 1u;
    if (b.obs_stale && obs) { for_each_env(b, [&](int64_t i) { obs_one<2>(b, i, obs); }, [&](int64_t i) { obs_one<4>(b, i, obs); }); b.obs_stale = false; }
}

// _TreasureGameImpl.init_with_state (impl:447-481) for every env; states [n][obs_dim] float64, -99 keeps the current value
void hostdev_init_with_state(void *h, const double *states) {
    Batch &b = *static_cast<Batch *>(h);
    if (b.B.tape) for_each_env(b, [&](int64_t i) { init_with_state_one<true, 2>(b, i, states); }, [&](int64_t i) { init_with_state_one<true, 4>(b, i, states); });
    else for_each_env(b, [&](int64_t i) { init_with_state_one<false, 2>(b, i, states); }, [&](int64_t i) { init_with_state_one<false, 4>(b, i, states); });
    b.obs_stale = true;
}

// one _TreasureGameImpl.step(action) (impl:290-359) for every env, like tg_primitive_step: every row is written
void hostdev_primitive_step(void *h, const int32_t *actions, float *obs, float *reward, uint8_t *done) {
    Batch &b = *static_cast<Batch *>(h);
    const uint32_t t = b.counter;
    if (b.B.tape) for_each_env(b, [&](int64_t i) { primitive_one<true, 2>(b, i, actions[i], t, obs, reward, done); }, [&](int64_t i) { primitive_one<true, 4>(b, i, actions[i], t, obs, reward, done); });
    else for_each_env(b, [&](int64_t i) { primitive_one<false, 2>(b, i, actions[i], t, obs, reward, done); }, [&](int64_t i) { primitive_one<false, 4>(b, i, actions[i], t, obs, reward, done); });
    b.counter = t + 1u;
    b.obs_stale = obs == nullptr;
}

// unpacked state, the layout of tg_get_state (strides TG_MAX_*): pos[n][2], misc[n][4] = facing, ticker, total_actions,
// draws; doors[n][6], handles[n][4], bolts[n][3] bytes; angles[n][4]; items[n][4][2]; bag[n][4] (-1 = empty); sticky[n]
void hostdev_state(void *h, int32_t *pos, int32_t *misc, uint8_t *doors, uint8_t *handles, uint8_t *bolts, double *angles,
                   int32_t *items, int32_t *bag, uint8_t *sticky) {
    Batch &b = *static_cast<Batch *>(h);
    for (int64_t i = 0; i < b.n; i++) {
        const uint4 c = b.core[i], a = b.acct[i];
        const uint32_t f = c.y;
        pos[i * 2] = core_px(c.x); pos[i * 2 + 1] = hi16(c.x);
        misc[i * 4] = f & 1u; misc[i * 4 + 1] = ticker(f); misc[i * 4 + 2] = (int)a.w; misc[i * 4 + 3] = (int)a.x;
        for (int j = 0; j < TG_MAX_DOORS; j++) doors[i * TG_MAX_DOORS + j] = (f >> (F_DOORS + j)) & 1u;
        for (int j = 0; j < TG_MAX_HANDLES; j++) handles[i * TG_MAX_HANDLES + j] = (f >> (F_HANDLES + j)) & 1u;
        for (int j = 0; j < TG_MAX_BOLTS; j++) bolts[i * TG_MAX_BOLTS + j] = (f >> (F_BOLTS + j)) & 1u;
        for (int j = 0; j < TG_MAX_HANDLES; j++) angles[i * TG_MAX_HANDLES + j] = b.angles[(size_t)j * b.n + i];
        uint32_t it[4] = {c.z, c.w, 0u, 0u};
        if (b.ni > 2) { it[2] = b.items23[i].x; it[3] = b.items23[i].y; }
        for (int j = 0; j < TG_MAX_ITEMS; j++) { items[(i * TG_MAX_ITEMS + j) * 2] = lo16(it[j]); items[(i * TG_MAX_ITEMS + j) * 2 + 1] = hi16(it[j]); }
        const int len = bag_len(f);
        for (int j = 0; j < TG_MAX_ITEMS; j++) bag[i * TG_MAX_ITEMS + j] = (j < len) ? (int)((f >> (F_BAGORD + 2 * j)) & 3u) : -1;
        sticky[i] = (uint8_t)core_sticky(c.x);
    }
}

void hostdev_mask(void *h, uint8_t *mask9) {                                         // tg_mask_kernel
    Batch &b = *static_cast<Batch *>(h);
    for (int64_t i = 0; i < b.n; i++)
        for (int k = 0; k < TG_NUM_OPTIONS; k++) mask9[i * TG_NUM_OPTIONS + k] = (uint8_t)((b.plan[i] >> k) & 1u);
}

// error flag and draw index of every env (the oracle's `error` / `draws`)
void hostdev_flags(void *h, uint8_t *error, uint32_t *draws) {
    Batch &b = *static_cast<Batch *>(h);
    for (int64_t i = 0; i < b.n; i++) { error[i] = (uint8_t)((b.core[i].y >> F_ERROR) & 1u); draws[i] = b.acct[i].x; }
}

}  // extern "C"
