// Dynamics kernels: step (option execution + reward + done + auto-reset + stats),
// reset, available-mask, state get/set.  One thread per environment; level blobs are
// staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) per CTA.
#include <cuda_runtime.h>
#include <cstdlib>
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

#ifndef TG_STEP_THREADS
#define TG_STEP_THREADS 256
#endif
constexpr int STEP_THREADS = TG_STEP_THREADS;
constexpr int AUX_THREADS = 128;       // reset / mask kernels
constexpr int NBUCKET = 64;          // length classes for the in-tile sort (bucket 0 = longest, 63 = not runnable)

// debug instrumentation (tg_debug_phase_buffer): thread 0 of every CTA stamps the phase boundaries
__device__ __forceinline__ void phase_stamp(const BatchView &B, int slot) {
    if (B.phase_ts && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        B.phase_ts[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

// info word per env of the tile: bit 0 runnable, bit 1 reference-would-raise, bits 2-7 target column + 8,
// bits 8-13 bucket, bits 14-17 option id (9 = not an option id), bit 18 runnable (kept by the later format)
__device__ __forceinline__ uint32_t pack_info(bool ran, bool err, int tcx, int bucket, int a) {
    return (ran ? 1u | (1u << 18) : 0u) | (err ? 2u : 0u) | ((uint32_t)((tcx + 8) & 63) << 2) | ((uint32_t)bucket << 8) | ((uint32_t)a << 14);
}

// Reset of an env whose option did not run (time limit): out of line and through memory, so that the common
// path of phase 4 carries neither the RNG state nor the Box-Muller code.  The caller has stored acct.
// `drawn` = uniforms the env has already consumed in this call (its option ran): the Philox blocks of a call are
// numbered from the call's first draw (draw_w), so the reset continues that numbering.
template <bool TAPE, int NI>
__device__ __noinline__ void reset_in_memory(const BatchView &B, const LevelBlob *L, int64_t i, uint32_t drawn) {
    Env<NI> e;
    uint4 acct;
    load_env(e, B, i, acct);
    e.d0 = e.draws - drawn;
    if (!TAPE && (drawn & 3u)) philox4x32_10(e.d0, drawn >> 2, e.id_lo, e.id_hi, e.key0, e.key1, e.w0, e.w1, e.w2, e.w3);
    reset_env<TAPE>(e, *L);
    store_env(e, B, i, acct);
}

// exclusive scan of hist[0 .. 63] by warp 0 (two entries per lane)
__device__ __forceinline__ void scan64(int *hist, int lane) {
    const int v0 = hist[2 * lane], v1 = hist[2 * lane + 1];
    int incl = v0 + v1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
    const int excl = incl - (v0 + v1);
    hist[2 * lane] = excl; hist[2 * lane + 1] = excl + v0;
}

// ---------------------------------------------------------------------------
// tg_step_kernel: one CTA owns a tile of `tile` (<= TILE) consecutive environments.
//   phase 0  counting sort of the tile by option id -> perm0[]: the classification below is a 9-way switch
//            on the option, and with i.i.d. actions a warp in index order runs every case with 3-4 lanes
//            (measured: can_run + length estimate were 30 % of the kernel's warp instructions at 2-6 active lanes)
//   phase 1  in option order: evaluate can_run of the chosen option (+ target column), estimate its
//            length in ticks; histogram the (code path, length) classes
//   phase 2  counting sort of the runnable envs by class -> perm[]; their 32-env chunks ordered longest first
//   phase 3  warps pull chunks from a shared counter: first the runnable chunks (a lane runs its env's option to
//            termination and puts the state back), then every index-order chunk for the envs whose option did not
//            run: time-limit / auto-reset / outputs with contiguous loads and full-line stores
//   phase 5  (after a barrier) the envs that ran: reward / done / time-limit / auto-reset / outputs
// Sorting puts the ~10-20 % runnable envs of a tile into a few full warps of similar length
// instead of leaving 1-2 busy lanes in every warp (measured SIMT efficiency before: 2/32 lanes).
// Results do not depend on the order: every env owns its RNG stream and state.
// ---------------------------------------------------------------------------
#ifndef TG_STEP_MIN_BLOCKS
#define TG_STEP_MIN_BLOCKS 3      // 80 registers, no spills, 3 CTAs per SM with 2364-env tiles: 9.9 G env-steps/s; 64 registers x 4 CTAs with
                                  // 1772-env tiles spilled 132 bytes: 9.2 G (round 1, v10 kernel; the v4 kernel had preferred 64 registers)
#endif
template <bool TAPE, int NI, int TILE>
__global__ void __launch_bounds__(STEP_THREADS, TG_STEP_MIN_BLOCKS)
tg_step_kernel(const __grid_constant__ BatchView B, int tile, const int32_t *__restrict__ actions, float *__restrict__ obs,
               float *__restrict__ reward, uint8_t *__restrict__ done_out, uint8_t *__restrict__ ran_out,
               uint16_t *__restrict__ avail_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    __shared__ uint64_t bar;
    __shared__ int sh_stats[8];
    __shared__ int hist[NBUCKET];
    __shared__ int next_chunk;
    __shared__ uint32_t rec_base;
    __shared__ uint8_t ckey[TILE / 32], order[TILE / 32];   // length class of each runnable chunk; chunks longest first
    __shared__ uint32_t info[TILE];
    __shared__ uint16_t perm0[TILE];
    uint16_t *const perm = perm0;                            // the class order replaces the option order (dead after phase 1)
    const int tid = threadIdx.x, lane = tid & 31;
    phase_stamp(B, 0);
    if (tid < 8) sh_stats[tid] = 0;
    if (tid < NBUCKET) hist[tid] = 0;
    if (tid == 0) next_chunk = 0;
    stage_levels(levels, B.levels, B.n_levels, &bar);      // contains a __syncthreads()
    phase_stamp(B, 1);

    const int64_t base = B.r_begin + (int64_t)blockIdx.x * tile;
    const int count = (int)min((int64_t)tile, B.r_begin + B.r_count - base);

    // Global loads of one phase are issued in batches before their first use: a dependent access costs about 1 us
    // here (measured with tg_debug_phase_buffer), and one element per thread at a time left the SM waiting on it.
    // ---- phase 0: sort by option id -------------------------------------------
    for (int k0 = 0; k0 < count; k0 += 8 * STEP_THREADS) {
        int av[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { const int el = k0 + u * STEP_THREADS + tid; av[u] = (el < count) ? actions[base + el] : 0; }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int el = k0 + u * STEP_THREADS + tid;
            if (el < count) {
                const int ac = ((unsigned)av[u] < (unsigned)TG_NUM_OPTIONS) ? av[u] : TG_NUM_OPTIONS;   // tg:92 raises IndexError; here: not run
                info[el] = (uint32_t)ac;
                atomicAdd(&hist[ac], 1);
            }
        }
    }
    __syncthreads();
    if (tid < 32) scan64(hist, lane);
    __syncthreads();
    for (int el = tid; el < count; el += STEP_THREADS) perm0[atomicAdd(&hist[info[el]], 1)] = (uint16_t)el;
    __syncthreads();
    if (tid < NBUCKET) hist[tid] = 0;
    __syncthreads();
    phase_stamp(B, 2);

    // ---- phase 1: classify, in option order -------------------------------------
    for (int j0 = 0; j0 < count; j0 += 4 * STEP_THREADS) {
        uint4 cv[4];
        int elv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = j0 + u * STEP_THREADS + tid;
            elv[u] = (j < count) ? perm0[j] : -1;
            if (elv[u] >= 0) cv[u] = B.core[base + elv[u]];
        }
#pragma unroll 1
        for (int u = 0; u < 4; u++) {       // one copy of the classification code: pick the element with selects
            const int el = (u == 0) ? elv[0] : (u == 1) ? elv[1] : (u == 2) ? elv[2] : elv[3];
            if (el < 0) continue;
            const uint4 cu = (u == 0) ? cv[0] : (u == 1) ? cv[1] : (u == 2) ? cv[2] : cv[3];
            const int64_t i = base + el;
            const LevelBlob &L = levels[B.level_id ? B.level_id[i] : 0];
            Env<NI> e;
            load_core(e, B, i, cu);
            const int a = (int)info[el];
            int tcx; bool err;
            const bool ran = option_setup(e, L, a, tcx, err);
            // sort key: code-path class major (lanes of a warp then run the same policy and the same branch of
            // tick()), estimated length minor (longest first); 63 = not runnable
            int bucket = NBUCKET - 1;
            if (ran) {
                const int cls = (a <= TG_GO_RIGHT) ? 0 : (a <= TG_DOWN_LADDER) ? 1 : (a == TG_INTERACT) ? 4 : (a <= TG_DOWN_RIGHT) ? 2 : 3;
                bucket = cls * 12 + 11 - min(estimate_ticks(e, L, a, tcx) / 10, 11);
            }
            info[el] = pack_info(ran, err, tcx, bucket, a);
            atomicAdd(&hist[bucket], 1);
        }
    }
    __syncthreads();
    phase_stamp(B, 3);
    // ---- phase 2: exclusive scan of the 64 class counts (warp 0), scatter the runnable envs, then order their
    // 32-env chunks longest first across classes (a chunk's first env is its longest: length is the minor key) ----
    if (tid < 32) scan64(hist, lane);
    __syncthreads();
    for (int el = tid; el < count; el += STEP_THREADS) {
        const uint32_t inf = info[el];
        if (inf & 1u) perm[atomicAdd(&hist[(inf >> 8) & 63], 1)] = (uint16_t)el;
    }
    const int n_run = hist[NBUCKET - 1];                    // exclusive scan: start of bucket 63 (not runnable) = runnable envs
    const int nrc = (n_run + 31) >> 5;
    __syncthreads();
    if (tid == STEP_THREADS - 1 && B.sp_count) rec_base = atomicAdd(B.sp_count, (uint32_t)n_run);    // sparse outputs: this tile's block of records
    if (tid < 32) {
        for (int c = lane; c < nrc; c += 32) ckey[c] = (uint8_t)(11 - ((info[perm[c * 32]] >> 8) & 63u) % 12u);
        __syncwarp();
        for (int c = lane; c < nrc; c += 32) {
            const int k = ckey[c];
            int r = 0;
            for (int j = 0; j < nrc; j++) { const int kj = ckey[j]; r += (kj > k || (kj == k && j < c)) ? 1 : 0; }
            order[r] = (uint8_t)c;
        }
    }
    __syncthreads();
    phase_stamp(B, 4);

    uint32_t st_cnt = 0;                                    // errors | episodes << 8 | successes << 16 | ran << 24
    int st_ticks = 0, st_ret = 0, st_epsteps = 0;
    const int od = B.obs_dim;
    float *stage = reinterpret_cast<float *>(smem_raw + (((size_t)B.n_levels * sizeof(LevelBlob) + 15) & ~(size_t)15))
                   + (size_t)(tid >> 5) * 32 * od;                             // this warp's [32][obs_dim] rows
    const bool vec_ok = obs && ((reinterpret_cast<uintptr_t>(obs + base * od) & 15u) == 0);    // this tile's rows start on 16 bytes

    // ---- phase 3: one queue of 32-env chunks from a shared counter.  First the runnable envs (class-sorted chunks,
    // longest first): a lane runs its env's option to termination and puts the state back; info[el] becomes bits
    // 0-12 ticks, bit 13 env newly flagged, bits 14-17 option id, bit 18 set, bits 19-31 uniforms drawn.  Then every
    // index-order chunk, for the envs whose option did not run (so this work overlaps the long options above):
    // time-limit / auto-reset (rare: out of line, through memory) / outputs with contiguous loads and stores.
    // Observation rows are built in shared memory and leave with full-line stores; the rows of the envs that ran
    // are rewritten by phase 5.
    const int nchunks = nrc + ((count + 31) >> 5);
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&next_chunk, 1);
        q = __shfl_sync(0xFFFFFFFFu, q, 0);
        if (q >= nchunks) break;
        if (q < nrc) {
            const int j = order[q] * 32 + lane;
            if (j < n_run) {
                const int el = perm[j];
                const uint32_t inf = info[el];
                const int64_t i = base + el;
                const LevelBlob &L = levels[B.level_id ? B.level_id[i] : 0];
                Env<NI> e;
                uint4 acct;
                load_env(e, B, i, acct);
                const int a = (int)((inf >> 14) & 15u);
                const uint32_t err0 = e.flags & (1u << F_ERROR);
                const int n = run_option_to_end<TAPE>(e, L, a, (int)((inf >> 2) & 63u) - 8);
                store_env(e, B, i, acct);
                const uint32_t newerr = ((e.flags & (1u << F_ERROR)) && !err0) ? 1u : 0u;
                info[el] = (uint32_t)n | (newerr << 13) | ((uint32_t)a << 14) | (1u << 18) | ((e.draws - e.d0) << 19);
            }
            continue;
        }
        const int el0 = (q - nrc) * 32, el = el0 + lane;
        const int rows = min(32, count - el0);
        if (el < count && !((info[el] >> 18) & 1u)) {
            const uint32_t inf = info[el];
            const int64_t i = base + el;
            const int lid = B.level_id ? B.level_id[i] : 0;
            const LevelBlob &L = levels[lid];
            Env<NI> e;
            uint4 cv = B.core[i];
            uint4 acct = B.acct[i];
            load_core(e, B, i, cv);
            e.angles = B.angles + i; e.n = B.n;
            if ((inf & 2u) && !(e.flags & (1u << F_ERROR))) {                // the reference would raise (target None)
                e.flags |= 1u << F_ERROR; st_cnt += 1u;
                cv.y = e.flags; B.core[i] = cv;
            }
            acct.z += 1u;
            const bool term = is_done(e, L);
            const bool trunc = B.max_steps > 0 && acct.z >= (uint32_t)B.max_steps;
            const int d = (term ? TG_DONE_TERMINATED : 0) | (trunc ? TG_DONE_TRUNCATED : 0);
            if (d) {
                st_cnt += (1u << 8) + (term ? 1u << 16 : 0u); st_ret += (int)acct.y; st_epsteps += (int)acct.z;
                if (B.auto_reset) {
                    acct.y = 0; acct.z = 0;
                    B.acct[i] = acct;
                    reset_in_memory<TAPE, NI>(B, &L, i, 0u);
                    acct = B.acct[i];
                    load_core(e, B, i);
                }
            }
            B.acct[i] = acct;
            if (B.sp_count) {                                                // sparse outputs: only an env that was reset (or stays done) reports
                if (d) {
                    uint32_t *rec = B.sp_recs + (size_t)atomicAdd(B.sp_count, 1u) * B.sp_words;
                    rec[0] = (uint32_t)i; rec[1] = 0u; rec[2] = (uint32_t)d;
                    write_obs(e, L, B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, reinterpret_cast<float *>(rec + 3), od);
                }
            }
            if (obs) write_obs(e, L, B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, stage + lane * od, od);
            if (reward) reward[i] = 0.0f;
            if (done_out) done_out[i] = (uint8_t)d;
            if (ran_out) ran_out[i] = (uint8_t)0;
            if (avail_out) avail_out[i] = (uint16_t)available_bits(e, L);
        }
        if (obs) {
            __syncwarp();
            float *dst = obs + (base + el0) * od;
            const int nf = rows * od;
            if (vec_ok && (nf & 3) == 0) {
                for (int k = lane; k < (nf >> 2); k += 32) reinterpret_cast<float4 *>(dst)[k] = reinterpret_cast<const float4 *>(stage)[k];
            } else {
                for (int k = lane; k < nf; k += 32) dst[k] = stage[k];
            }
            __syncwarp();
        }
    }
    __syncthreads();      // every option has run; the rows written above are ordered before the ones written below
    phase_stamp(B, 5);

    // ---- phase 5: the envs whose option ran: reward / done / time-limit / auto-reset / outputs ----
    for (int j = tid; j < n_run; j += STEP_THREADS) {
        const int el = perm[j];
        const uint32_t inf = info[el];
        const int64_t i = base + el;
        const int lid = B.level_id ? B.level_id[i] : 0;
        const LevelBlob &L = levels[lid];
        Env<NI> e;
        uint4 acct = B.acct[i];
        load_core(e, B, i);
        e.angles = B.angles + i; e.n = B.n;
        const int a = (int)((inf >> 14) & 15u);
        const int n = (int)(inf & 0x1FFFu);
        st_cnt += ((inf >> 13) & 1u) + (1u << 24);
        st_ticks += n;
        const int r = -n - ((a >= TG_JUMP_LEFT) ? 4 : 0);                    // impl:15-16: -1 per tick, JUMP tick -5
        acct.y = (uint32_t)((int)acct.y + r);
        acct.z += 1u;
        const bool term = is_done(e, L);
        const bool trunc = B.max_steps > 0 && acct.z >= (uint32_t)B.max_steps;
        const int d = (term ? TG_DONE_TERMINATED : 0) | (trunc ? TG_DONE_TRUNCATED : 0);
        if (d) {
            st_cnt += (1u << 8) + (term ? 1u << 16 : 0u); st_ret += (int)acct.y; st_epsteps += (int)acct.z;
            if (B.auto_reset) {
                acct.y = 0; acct.z = 0;
                B.acct[i] = acct;
                reset_in_memory<TAPE, NI>(B, &L, i, inf >> 19);
                acct = B.acct[i];
                load_core(e, B, i);
            }
        }
        B.acct[i] = acct;
        if (B.sp_count) {                                                    // sparse outputs: every env that ran reports
            uint32_t *rec = B.sp_recs + (size_t)(rec_base + (uint32_t)j) * B.sp_words;
            rec[0] = (uint32_t)i; rec[1] = __float_as_uint((float)r); rec[2] = (uint32_t)d | 256u;
            write_obs(e, L, B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, reinterpret_cast<float *>(rec + 3), od);
        }
        if (obs) write_obs(e, L, B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, obs + i * od, od);
        if (reward) reward[i] = (float)r;
        if (done_out) done_out[i] = (uint8_t)d;
        if (ran_out) ran_out[i] = (uint8_t)1;
        if (avail_out) avail_out[i] = (uint16_t)available_bits(e, L);
    }
    phase_stamp(B, 6);
    int st[8];
    st[ST_EPISODES] = (st_cnt >> 8) & 255; st[ST_SUCCESS] = (st_cnt >> 16) & 255; st[ST_RETURN] = st_ret; st[ST_EPSTEPS] = st_epsteps;
    st[ST_TICKS] = st_ticks; st[ST_RAN] = st_cnt >> 24; st[ST_ERRORS] = st_cnt & 255;
    st[ST_STEPS] = (tid == 0) ? count : 0;                  // every env of the tile takes exactly one gym step
    stats_accumulate(sh_stats, B.stats, st);
    phase_stamp(B, 7);
}

template <bool TAPE, int NI>
__global__ void __launch_bounds__(AUX_THREADS)
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
    if (obs) write_obs(e, L, B.obs_lut + (size_t)(B.level_id ? B.level_id[i] : 0) * 2 * OBS_LUT_N, obs + i * B.obs_dim, B.obs_dim);
}

template <int NI>
__global__ void __launch_bounds__(AUX_THREADS)
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

// One primitive action per env: _TreasureGameImpl.step(act) (impl:290-359) without the option layer.
// Action ids are _actions.py:7-13 (anything else falls through every branch like NOP); reward -1, JUMP -5.
// Episode accounting, done / truncation / auto-reset and statistics behave as in tg_step_kernel.
template <bool TAPE, int NI>
__global__ void __launch_bounds__(AUX_THREADS)
tg_primitive_kernel(BatchView B, const int32_t *__restrict__ actions, float *__restrict__ obs,
                    float *__restrict__ reward, uint8_t *__restrict__ done_out) {
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
        tick<TAPE>(e, L, a);
        const int r = (a == A_JUMP) ? -5 : -1;                               // impl:15-16,356-359
        acct.y = (uint32_t)((int)acct.y + r);
        acct.z += 1u;
        const bool term = is_done(e, L);
        const bool trunc = B.max_steps > 0 && acct.z >= (uint32_t)B.max_steps;
        const int d = (term ? TG_DONE_TERMINATED : 0) | (trunc ? TG_DONE_TRUNCATED : 0);
        st[ST_TICKS] = 1; st[ST_RAN] = 1; st[ST_STEPS] = 1;
        st[ST_ERRORS] = ((e.flags & (1u << F_ERROR)) && !err0) ? 1 : 0;
        if (d) {
            st[ST_EPISODES] = 1; st[ST_SUCCESS] = term; st[ST_RETURN] = (int)acct.y; st[ST_EPSTEPS] = (int)acct.z;
            if (B.auto_reset) { reset_env<TAPE>(e, L); acct.y = 0; acct.z = 0; }
        }
        store_env(e, B, i, acct);
        if (obs) write_obs(e, L, B.obs_lut + (size_t)(B.level_id ? B.level_id[i] : 0) * 2 * OBS_LUT_N, obs + i * B.obs_dim, B.obs_dim);
        if (reward) reward[i] = (float)r;
        if (done_out) done_out[i] = (uint8_t)d;
    }
    stats_accumulate(sh_stats, B.stats, st);
}

template <bool TAPE, int NI>
__global__ void __launch_bounds__(AUX_THREADS)
tg_init_with_state_kernel(BatchView B, const double *__restrict__ states, const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    __shared__ uint64_t bar;
    stage_levels(levels, B.levels, B.n_levels, &bar);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n || (mask && !mask[i])) return;
    const LevelBlob &L = levels[B.level_id ? B.level_id[i] : 0];
    Env<NI> e;
    uint4 acct;
    load_env(e, B, i, acct);
    init_with_state_env<TAPE>(e, L, states + i * B.obs_dim);
    store_env(e, B, i, acct);
}

// ---- state get / set (unpacked view, strides = TG_MAX_*) -------------------
__global__ void tg_get_state_kernel(BatchView B, tg_state_view v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    const uint4 c = B.core[i], a = B.acct[i];
    const uint32_t f = c.y;
    if (v.pos) { v.pos[i * 2] = core_px(c.x); v.pos[i * 2 + 1] = hi16(c.x); }
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
    if (v.acct) { v.acct[i * 3] = (int)a.y; v.acct[i * 3 + 1] = a.z; v.acct[i * 3 + 2] = ((f >> F_ERROR) & 1u) | (core_sticky(c.x) << 1); }
}

__global__ void tg_set_state_kernel(BatchView B, tg_state_view v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    uint4 c = B.core[i], a = B.acct[i];
    uint32_t f = c.y;
    if (v.pos) {   // keep the probe invariants of tg_device.cuh (pad_cell has no clamps)
        const LevelBlob &L = B.levels[B.level_id ? B.level_id[i] : 0];
        const int x = min(max(v.pos[i * 2], 0), L.cw * S - 1), y = min(max(v.pos[i * 2 + 1], -(S - 1)), L.ch * S - 1);
        c.x = pack_player(x, y, core_sticky(c.x));
    }
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
        f &= ~(0xFFFu << F_BAGLEN);
        int len = 0;
        for (int j = 0; j < TG_MAX_ITEMS; j++) {
            int it = v.bag[i * TG_MAX_ITEMS + j];
            if (it < 0 || it >= TG_MAX_ITEMS) break;
            f |= (uint32_t)it << (F_BAGORD + 2 * j);
            len++;
        }
        f |= (uint32_t)len << F_BAGLEN;
    }
    if (v.acct) {
        a.y = (uint32_t)(int)v.acct[i * 3]; a.z = (uint32_t)v.acct[i * 3 + 1];
        f = (f & ~(1u << F_ERROR)) | (((v.acct[i * 3 + 2] & 1) ? 1u : 0u) << F_ERROR);
        c.x = pack_player(core_px(c.x), hi16(c.x), (uint32_t)(v.acct[i * 3 + 2] >> 1));
    }
    c.y = f;
    B.core[i] = c; B.acct[i] = a;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
static inline size_t level_smem(const BatchView &B) { return (size_t)B.n_levels * sizeof(LevelBlob); }
// step kernel: level blobs + one [32][obs_dim] float staging area per warp for the coalesced observation rows
static inline size_t step_smem(const BatchView &B) {
    return ((level_smem(B) + 15) & ~(size_t)15) + (size_t)(STEP_THREADS / 32) * 32 * B.obs_dim * sizeof(float);
}

template <bool TAPE, int NI, int TILE>
static cudaError_t step_tile(const BatchView &B, int tile, const int32_t *a, float *obs, float *rew, uint8_t *done,
                             uint8_t *ran, uint16_t *avail, cudaStream_t s) {
    const size_t smem = step_smem(B);
    {   // static + dynamic shared memory can pass the default 48 KB (large tiles, many layouts): opt in once per size
        static size_t allowed = 0;
        if (smem > allowed) {
            cudaError_t r = cudaFuncSetAttribute(tg_step_kernel<TAPE, NI, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (r != cudaSuccess) return r;
            allowed = smem;
        }
    }
    tg_step_kernel<TAPE, NI, TILE><<<grid_for(B.r_count, tile), STEP_THREADS, step_smem(B), s>>>(B, tile, a, obs, rew, done, ran, avail);
    return cudaGetLastError();
}

template <bool TAPE, int NI>
static cudaError_t step_impl(const BatchView &B, int tile, const int32_t *a, float *obs, float *rew, uint8_t *done,
                             uint8_t *ran, uint16_t *avail, cudaStream_t s) {
    if (tile <= 256) return step_tile<TAPE, NI, 256>(B, tile, a, obs, rew, done, ran, avail, s);     // capacity of the shared arrays
    if (tile <= 1024) return step_tile<TAPE, NI, 1024>(B, tile, a, obs, rew, done, ran, avail, s);
    return step_tile<TAPE, NI, 4096>(B, tile, a, obs, rew, done, ran, avail, s);
}

// Tile size: large tiles sort better (more runnable envs per tile -> fuller warps); small tiles give more
// CTAs.  The grid is sized in whole "slots": 148 SMs x 3 resident CTAs = 444 CTAs run at once, so the tile is
// n / (444 * waves) rounded up (2364 envs for a 1,048,576-env step) -- with fixed 2048-env tiles and 4 CTAs per SM a
// step had 512 CTAs, 68 SMs held four of them and 80 SMs three.  Smaller batches aim for two, then one CTA per SM slot.  TG_STEP_TILE overrides.
int pick_step_tile(int64_t n) {
    static int forced = -1, slots = 0;
    if (forced < 0) { const char *v = getenv("TG_STEP_TILE"); forced = v ? atoi(v) : 0; }
    if (forced >= 32 && forced <= 4096) return (forced + 3) / 4 * 4;
    if (!slots) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = sms * TG_STEP_MIN_BLOCKS;
    }
    const int64_t cap = 4096;
    int64_t ctas;
    if (n >= (int64_t)slots * 512) ctas = (n + slots * cap - 1) / (slots * cap) * slots;      // whole waves of full occupancy
    else if (n >= (int64_t)slots * 128) ctas = slots / 2;                                      // two CTAs per SM
    else ctas = (n + 255) / 256;
    int64_t tile = (n + ctas - 1) / ctas;
    tile = (tile + 3) / 4 * 4;
    if (tile < 32) tile = 32;
    if (tile > 4096) tile = 4096;
    return (int)tile;
}

cudaError_t launch_step(const BatchView &B, int ni, const int32_t *a, float *obs, float *rew, uint8_t *done,
                        uint8_t *ran, uint16_t *avail, cudaStream_t s) {
    const bool tape = B.tape != nullptr;
    const int tile = pick_step_tile(B.r_count);
    if (ni <= 2) return tape ? step_impl<true, 2>(B, tile, a, obs, rew, done, ran, avail, s) : step_impl<false, 2>(B, tile, a, obs, rew, done, ran, avail, s);
    return tape ? step_impl<true, 4>(B, tile, a, obs, rew, done, ran, avail, s) : step_impl<false, 4>(B, tile, a, obs, rew, done, ran, avail, s);
}

cudaError_t launch_reset(const BatchView &B, int ni, const uint8_t *mask, float *obs, cudaStream_t s) {
    const unsigned g = grid_for(B.n, AUX_THREADS);
    const size_t sm = level_smem(B);
    const bool tape = B.tape != nullptr;
    if (ni <= 2) {
        if (tape) tg_reset_kernel<true, 2><<<g, AUX_THREADS, sm, s>>>(B, mask, obs);
        else tg_reset_kernel<false, 2><<<g, AUX_THREADS, sm, s>>>(B, mask, obs);
    } else {
        if (tape) tg_reset_kernel<true, 4><<<g, AUX_THREADS, sm, s>>>(B, mask, obs);
        else tg_reset_kernel<false, 4><<<g, AUX_THREADS, sm, s>>>(B, mask, obs);
    }
    return cudaGetLastError();
}

cudaError_t launch_primitive(const BatchView &B, int ni, const int32_t *a, float *obs, float *rew, uint8_t *done, cudaStream_t s) {
    const unsigned g = grid_for(B.n, AUX_THREADS);
    const size_t sm = level_smem(B);
    const bool tape = B.tape != nullptr;
    if (ni <= 2) {
        if (tape) tg_primitive_kernel<true, 2><<<g, AUX_THREADS, sm, s>>>(B, a, obs, rew, done);
        else tg_primitive_kernel<false, 2><<<g, AUX_THREADS, sm, s>>>(B, a, obs, rew, done);
    } else {
        if (tape) tg_primitive_kernel<true, 4><<<g, AUX_THREADS, sm, s>>>(B, a, obs, rew, done);
        else tg_primitive_kernel<false, 4><<<g, AUX_THREADS, sm, s>>>(B, a, obs, rew, done);
    }
    return cudaGetLastError();
}

cudaError_t launch_init_with_state(const BatchView &B, int ni, const double *states, const uint8_t *mask, cudaStream_t s) {
    const unsigned g = grid_for(B.n, AUX_THREADS);
    const size_t sm = level_smem(B);
    const bool tape = B.tape != nullptr;
    if (ni <= 2) {
        if (tape) tg_init_with_state_kernel<true, 2><<<g, AUX_THREADS, sm, s>>>(B, states, mask);
        else tg_init_with_state_kernel<false, 2><<<g, AUX_THREADS, sm, s>>>(B, states, mask);
    } else {
        if (tape) tg_init_with_state_kernel<true, 4><<<g, AUX_THREADS, sm, s>>>(B, states, mask);
        else tg_init_with_state_kernel<false, 4><<<g, AUX_THREADS, sm, s>>>(B, states, mask);
    }
    return cudaGetLastError();
}

cudaError_t launch_mask(const BatchView &B, int ni, uint8_t *mask, cudaStream_t s) {
    const unsigned g = grid_for(B.n, AUX_THREADS);
    if (ni <= 2) tg_mask_kernel<2><<<g, AUX_THREADS, level_smem(B), s>>>(B, mask);
    else tg_mask_kernel<4><<<g, AUX_THREADS, level_smem(B), s>>>(B, mask);
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
