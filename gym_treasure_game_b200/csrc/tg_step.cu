// Dynamics kernels: step (option execution + reward + done + auto-reset + stats),
// reset, available-mask, state get/set.  One thread per environment; level blobs are
// staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) per CTA.
#include <cuda_runtime.h>
#include "tg_device.cuh"
#include "tg_launch.h"

namespace tg {

// ---------------------------------------------------------------------------
// TMA bulk staging of the level blobs (global -> shared), SASS: UBLKCP + SYNCS
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_levels(LevelBlob *dst, const LevelBlob *src, int n_levels, uint64_t *bar) {
    const uint32_t bar_a = smem_u32(bar), dst_a = smem_u32(dst);
    const uint32_t bytes = (uint32_t)n_levels * (uint32_t)sizeof(LevelBlob);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_a), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
    }
    __syncthreads();     // barrier init visible to every waiter
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar_a) : "memory");
    } while (!ok);
}

// block-level accumulation of the 8 statistics: warp REDUX -> shared atomics -> 8 global atomics per CTA
__device__ __forceinline__ void stats_accumulate(int *sh, unsigned long long *gstats, const int (&v)[8]) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int s = __reduce_add_sync(0xFFFFFFFFu, v[j]);
        if (lane == 0 && s != 0) atomicAdd(&sh[j], s);
    }
    __syncthreads();
    if (threadIdx.x < 8 && sh[threadIdx.x] != 0)
        atomicAdd(&gstats[threadIdx.x], (unsigned long long)(long long)sh[threadIdx.x]);
}

constexpr int STEP_THREADS = 128;

template <bool TAPE, int NI>
__global__ void __launch_bounds__(STEP_THREADS)
tg_step_kernel(BatchView B, const int32_t *__restrict__ actions, float *__restrict__ obs,
               float *__restrict__ reward, uint8_t *__restrict__ done_out, uint8_t *__restrict__ ran_out,
               uint16_t *__restrict__ avail_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    __shared__ uint64_t bar;
    __shared__ int sh_stats[8];
    if (threadIdx.x < 8) sh_stats[threadIdx.x] = 0;
    stage_levels(levels, B.levels, B.n_levels, &bar);

    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (i < B.n) {
        const LevelBlob &L = levels[B.level_id ? B.level_id[i] : 0];
        Env<NI> e;
        uint4 acct;
        load_env(e, B, i, acct);
        const int a = actions[i];
        const uint32_t err0 = e.flags & (1u << F_ERROR);

        const int n = run_option<TAPE>(e, L, a);
        const int r = n ? -n - ((a >= TG_JUMP_LEFT) ? 4 : 0) : 0;       // impl:15-16: -1 per tick, JUMP tick -5
        acct.y = (uint32_t)((int)acct.y + r);
        acct.z += 1u;
        const bool term = is_done(e, L);
        const bool trunc = B.max_steps > 0 && acct.z >= (uint32_t)B.max_steps;
        const int d = (term ? TG_DONE_TERMINATED : 0) | (trunc ? TG_DONE_TRUNCATED : 0);
        st[ST_TICKS] = n; st[ST_RAN] = n > 0; st[ST_STEPS] = 1;
        st[ST_ERRORS] = ((e.flags & (1u << F_ERROR)) && !err0) ? 1 : 0;
        if (d) {
            st[ST_EPISODES] = 1; st[ST_SUCCESS] = term; st[ST_RETURN] = (int)acct.y; st[ST_EPSTEPS] = (int)acct.z;
            if (B.auto_reset) { reset_env<TAPE>(e, L); acct.y = 0; acct.z = 0; }
        }
        store_env(e, B, i, acct);
        if (obs) write_obs(e, L, obs + i * B.obs_dim, B.obs_dim);
        if (reward) reward[i] = (float)r;
        if (done_out) done_out[i] = (uint8_t)d;
        if (ran_out) ran_out[i] = (uint8_t)(n > 0);
        if (avail_out) avail_out[i] = (uint16_t)available_bits(e, L);
    }
    stats_accumulate(sh_stats, B.stats, st);
}

template <bool TAPE, int NI>
__global__ void __launch_bounds__(STEP_THREADS)
tg_reset_kernel(BatchView B, const uint8_t *__restrict__ mask, float *__restrict__ obs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    __shared__ uint64_t bar;
    stage_levels(levels, B.levels, B.n_levels, &bar);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    const LevelBlob &L = levels[B.level_id ? B.level_id[i] : 0];
    Env<NI> e;
    uint4 acct;
    load_env(e, B, i, acct);
    if (!mask || mask[i]) {
        e.flags &= ~(1u << F_ERROR);
        reset_env<TAPE>(e, L);
        acct.y = 0; acct.z = 0;
        store_env(e, B, i, acct);
    }
    if (obs) write_obs(e, L, obs + i * B.obs_dim, B.obs_dim);
}

template <int NI>
__global__ void __launch_bounds__(STEP_THREADS)
tg_mask_kernel(BatchView B, uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    __shared__ uint64_t bar;
    stage_levels(levels, B.levels, B.n_levels, &bar);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    const LevelBlob &L = levels[B.level_id ? B.level_id[i] : 0];
    Env<NI> e;
    uint4 acct;
    load_env(e, B, i, acct);
    const uint32_t m = available_bits(e, L);
#pragma unroll
    for (int k = 0; k < TG_NUM_OPTIONS; k++) mask[i * TG_NUM_OPTIONS + k] = (m >> k) & 1u;
}

// ---- state get / set (unpacked view, strides = TG_MAX_*) -------------------
__global__ void tg_get_state_kernel(BatchView B, tg_state_view v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    const uint4 c = B.core[i], a = B.acct[i];
    const uint32_t f = c.y;
    if (v.pos) { v.pos[i * 2] = lo16(c.x); v.pos[i * 2 + 1] = hi16(c.x); }
    if (v.misc) { v.misc[i * 4] = f & 1u; v.misc[i * 4 + 1] = ticker(f); v.misc[i * 4 + 2] = (int)a.w; v.misc[i * 4 + 3] = (int)a.x; }
    if (v.doors) for (int j = 0; j < TG_MAX_DOORS; j++) v.doors[i * TG_MAX_DOORS + j] = (f >> (F_DOORS + j)) & 1u;
    if (v.handles) for (int j = 0; j < TG_MAX_HANDLES; j++) v.handles[i * TG_MAX_HANDLES + j] = (f >> (F_HANDLES + j)) & 1u;
    if (v.bolts) for (int j = 0; j < TG_MAX_BOLTS; j++) v.bolts[i * TG_MAX_BOLTS + j] = (f >> (F_BOLTS + j)) & 1u;
    if (v.angles) for (int j = 0; j < TG_MAX_HANDLES; j++) v.angles[i * TG_MAX_HANDLES + j] = B.angles[(int64_t)j * B.n + i];
    uint32_t it[4] = {c.z, c.w, 0u, 0u};
    if (B.items23) { uint2 h = B.items23[i]; it[2] = h.x; it[3] = h.y; }
    if (v.items) for (int j = 0; j < TG_MAX_ITEMS; j++) { v.items[(i * TG_MAX_ITEMS + j) * 2] = lo16(it[j]); v.items[(i * TG_MAX_ITEMS + j) * 2 + 1] = hi16(it[j]); }
    if (v.bag) {
        const int len = bag_len(f);
        for (int j = 0; j < TG_MAX_ITEMS; j++) v.bag[i * TG_MAX_ITEMS + j] = (j < len) ? (int)((f >> (F_BAGORD + 2 * j)) & 3u) : -1;
    }
    if (v.acct) { v.acct[i * 3] = (int)a.y; v.acct[i * 3 + 1] = a.z; v.acct[i * 3 + 2] = (f >> F_ERROR) & 1u; }
}

__global__ void tg_set_state_kernel(BatchView B, tg_state_view v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    uint4 c = B.core[i], a = B.acct[i];
    uint32_t f = c.y;
    if (v.pos) c.x = pack_xy(v.pos[i * 2], v.pos[i * 2 + 1]);
    if (v.misc) {
        f = (f & ~1u) | (v.misc[i * 4] ? 1u : 0u);
        f = set_ticker(f, v.misc[i * 4 + 1] & 31);
        a.w = (uint32_t)v.misc[i * 4 + 2]; a.x = (uint32_t)v.misc[i * 4 + 3];
    }
    if (v.doors) for (int j = 0; j < TG_MAX_DOORS; j++) f = (f & ~(1u << (F_DOORS + j))) | ((v.doors[i * TG_MAX_DOORS + j] ? 1u : 0u) << (F_DOORS + j));
    if (v.handles) for (int j = 0; j < TG_MAX_HANDLES; j++) f = (f & ~(1u << (F_HANDLES + j))) | ((v.handles[i * TG_MAX_HANDLES + j] ? 1u : 0u) << (F_HANDLES + j));
    if (v.bolts) for (int j = 0; j < TG_MAX_BOLTS; j++) f = (f & ~(1u << (F_BOLTS + j))) | ((v.bolts[i * TG_MAX_BOLTS + j] ? 1u : 0u) << (F_BOLTS + j));
    if (v.angles) for (int j = 0; j < TG_MAX_HANDLES; j++) B.angles[(int64_t)j * B.n + i] = v.angles[i * TG_MAX_HANDLES + j];
    if (v.items) {
        c.z = pack_xy(v.items[(i * TG_MAX_ITEMS + 0) * 2], v.items[(i * TG_MAX_ITEMS + 0) * 2 + 1]);
        c.w = pack_xy(v.items[(i * TG_MAX_ITEMS + 1) * 2], v.items[(i * TG_MAX_ITEMS + 1) * 2 + 1]);
        if (B.items23) {
            uint2 h;
            h.x = pack_xy(v.items[(i * TG_MAX_ITEMS + 2) * 2], v.items[(i * TG_MAX_ITEMS + 2) * 2 + 1]);
            h.y = pack_xy(v.items[(i * TG_MAX_ITEMS + 3) * 2], v.items[(i * TG_MAX_ITEMS + 3) * 2 + 1]);
            B.items23[i] = h;
        }
    }
    if (v.bag) {
        f &= ~(0xFFFu << F_INBAG);
        for (int j = 0; j < TG_MAX_ITEMS; j++) {
            int it = v.bag[i * TG_MAX_ITEMS + j];
            if (it < 0 || it >= TG_MAX_ITEMS) break;
            f |= 1u << (F_INBAG + it);
            f |= (uint32_t)it << (F_BAGORD + 2 * j);
        }
    }
    if (v.acct) {
        a.y = (uint32_t)(int)v.acct[i * 3]; a.z = (uint32_t)v.acct[i * 3 + 1];
        f = (f & ~(1u << F_ERROR)) | ((v.acct[i * 3 + 2] ? 1u : 0u) << F_ERROR);
    }
    c.y = f;
    B.core[i] = c; B.acct[i] = a;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
static inline size_t level_smem(const BatchView &B) { return (size_t)B.n_levels * sizeof(LevelBlob); }

template <bool TAPE, int NI>
static cudaError_t step_impl(const BatchView &B, const int32_t *a, float *obs, float *rew, uint8_t *done,
                             uint8_t *ran, uint16_t *avail, cudaStream_t s) {
    tg_step_kernel<TAPE, NI><<<grid_for(B.n, STEP_THREADS), STEP_THREADS, level_smem(B), s>>>(B, a, obs, rew, done, ran, avail);
    return cudaGetLastError();
}

cudaError_t launch_step(const BatchView &B, int ni, const int32_t *a, float *obs, float *rew, uint8_t *done,
                        uint8_t *ran, uint16_t *avail, cudaStream_t s) {
    const bool tape = B.tape != nullptr;
    if (ni <= 2) return tape ? step_impl<true, 2>(B, a, obs, rew, done, ran, avail, s) : step_impl<false, 2>(B, a, obs, rew, done, ran, avail, s);
    return tape ? step_impl<true, 4>(B, a, obs, rew, done, ran, avail, s) : step_impl<false, 4>(B, a, obs, rew, done, ran, avail, s);
}

cudaError_t launch_reset(const BatchView &B, int ni, const uint8_t *mask, float *obs, cudaStream_t s) {
    const unsigned g = grid_for(B.n, STEP_THREADS);
    const size_t sm = level_smem(B);
    const bool tape = B.tape != nullptr;
    if (ni <= 2) {
        if (tape) tg_reset_kernel<true, 2><<<g, STEP_THREADS, sm, s>>>(B, mask, obs);
        else tg_reset_kernel<false, 2><<<g, STEP_THREADS, sm, s>>>(B, mask, obs);
    } else {
        if (tape) tg_reset_kernel<true, 4><<<g, STEP_THREADS, sm, s>>>(B, mask, obs);
        else tg_reset_kernel<false, 4><<<g, STEP_THREADS, sm, s>>>(B, mask, obs);
    }
    return cudaGetLastError();
}

cudaError_t launch_mask(const BatchView &B, int ni, uint8_t *mask, cudaStream_t s) {
    const unsigned g = grid_for(B.n, STEP_THREADS);
    if (ni <= 2) tg_mask_kernel<2><<<g, STEP_THREADS, level_smem(B), s>>>(B, mask);
    else tg_mask_kernel<4><<<g, STEP_THREADS, level_smem(B), s>>>(B, mask);
    return cudaGetLastError();
}

cudaError_t launch_get_state(const BatchView &B, const tg_state_view &v, cudaStream_t s) {
    tg_get_state_kernel<<<grid_for(B.n, 256), 256, 0, s>>>(B, v);
    return cudaGetLastError();
}
cudaError_t launch_set_state(const BatchView &B, const tg_state_view &v, cudaStream_t s) {
    tg_set_state_kernel<<<grid_for(B.n, 256), 256, 0, s>>>(B, v);
    return cudaGetLastError();
}

}  // namespace tg
