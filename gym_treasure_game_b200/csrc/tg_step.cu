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
#define TG_STEP_THREADS 384      // 12 warps x 2 CTAs per SM at 80 registers (256 x 3: 103 -> 99 us per 1,048,576-env step; 512 x 2 at 64 registers spills: 101 us)
#endif
constexpr int STEP_THREADS = TG_STEP_THREADS;
constexpr int AUX_THREADS = 128;       // reset / mask kernels
constexpr int NBUCKET = 64;          // length classes for the in-tile sort (bucket 0 = longest, 63 = not runnable)

// debug instrumentation (tg_debug_phase_buffer): thread 0 of every CTA stamps the phase boundaries.  The timer read
// is made to depend on a shared-memory load: BAR.SYNC.DEFER_BLOCKING lets a warp run ahead of the barrier until its next
// memory access, so a bare special-register read right after __syncthreads() is taken before the barrier completes.
__device__ __forceinline__ void phase_stamp(const BatchView &B, int slot, const volatile int *sh) {
    if (B.phase_ts && threadIdx.x == 0) {
        unsigned long long t = 0;
        if (*sh != -0x7fffffff) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory");
        B.phase_ts[(size_t)blockIdx.x * 8 + slot] = t;
    }
}

// The last CTA of a launch to finish (every CTA has read step_counter[0] by then) advances the batch's gym-step
// counter when the launch is the last one of its API call, and re-arms the ticket.
__device__ __forceinline__ void finish_launch(const BatchView &B) {
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(&B.step_counter[1], 1u) == total - 1u) {
            B.step_counter[1] = 0u;
            if (B.advance) B.step_counter[0] += 1u;
        }
    }
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

// The interact option (opts:442-460: one INTERACT tick, impl:321-329) through memory and out of line: lever flip,
// trigger graph, handle angles and the key drop are rare (2.6 % of the steps under random actions) and large; kept out
// of the option lanes' instruction stream.  Returns the uniforms drawn.
template <bool TAPE, int NI>
__device__ __noinline__ uint32_t interact_option_mem(const BatchView &B, const LevelBlob *Lp, int64_t i) {
    Env<NI> e;
    uint4 acct;
    load_env(e, B, i, acct);
    tick<TAPE, NI, true>(e, *Lp, A_INTERACT);
    store_env(e, B, i, acct);
    return e.draws - e.d0;
}

// acct.z of an env that ran and whose episode ended, between its option lane and the reset pass:
// uniforms drawn in this call (16 bits) | primitive ticks (13) | jump option (1) | done bits (2)
__device__ __forceinline__ uint32_t pack_pending(uint32_t drawn, int n, bool jump, uint32_t d) {
    return min(drawn, 0xFFFFu) | ((uint32_t)n << 16) | ((jump ? 1u : 0u) << 29) | (d << 30);
}

// ---------------------------------------------------------------------------
// tg_step_kernel: one CTA owns a tile of `tile` consecutive environments (shared arrays sized for `cap` >= tile).
//   phase A  index order, vector loads / stores: action + plan word + episode start of every env.  An env whose
//            option cannot run (80 % under random actions; the reference returns None and leaves the state untouched,
//            opt:22-23) is finished here: reward 0, ran 0, done (terminated from the plan, truncated from the episode
//            start) -- 16 bytes read and 6 written, its state record is never loaded and its observation row is
//            not rewritten (see tg_step in treasure_b200.h: rows only change when the env's state does).  A runnable
//            env gets its sort bucket (code-path class major, length class from the plan minor) and its rank in the
//            bucket (one shared-memory atomic).  Idle envs whose episode ended go on the reset list.
//   sort     exclusive scan of the 64 bucket counts, scatter by rank -> perm[]; 32-env chunks ordered longest first
//   phase B  warps pull chunks from a shared counter: a lane loads its env, runs the option to termination
//            (run_option_to_end = opt:20-36 + impl:290-359) and finishes the gym step in place: reward, done
//            (tg:95), time limit, new plan, state, observation row, outputs; an env whose episode ended goes on
//            the reset list instead.
//   phase C  (after a barrier) the reset list, full warps: reset_game (impl:55-73), plan, state, observation row.
// Sorting puts the ~20 % runnable envs of a tile into full warps that run the same code path for a similar number
// of ticks (one thread per env in index order: 2 of 32 lanes active, profiles/r01_step_v1_ncu.txt).  Results do not
// depend on the order: every env owns its RNG stream and state.
// ---------------------------------------------------------------------------
#ifndef TG_STEP_MIN_BLOCKS
#define TG_STEP_MIN_BLOCKS 2
#endif
template <bool TAPE, int NI>
__global__ void __launch_bounds__(STEP_THREADS, TG_STEP_MIN_BLOCKS)
tg_step_kernel(const __grid_constant__ BatchView B, int tile, int cap, const int32_t *__restrict__ actions, float *__restrict__ obs,
               float *__restrict__ reward, uint8_t *__restrict__ done_out, uint8_t *__restrict__ ran_out,
               uint16_t *__restrict__ avail_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    uint8_t *code = smem_raw + (((size_t)B.n_levels * sizeof(LevelBlob) + 15) & ~(size_t)15);   // [cap] sort bucket, 255 = idle
    uint16_t *rank = reinterpret_cast<uint16_t *>(code + cap);   // [cap] rank inside the bucket; later: sorted position of an env to reset
    uint16_t *perm = rank + cap;                                 // [cap] runnable envs, sorted
    uint16_t *rlist = perm + cap;                                // [cap] envs to reset: el | 0x8000 when its option ran
    __shared__ uint64_t bar;
    __shared__ int sh_stats[8];
    __shared__ int hist[NBUCKET];
    __shared__ int next_chunk, n_reset;
    __shared__ uint32_t rec_base;
    __shared__ uint8_t ckey[136], order[136], clen[136];    // per chunk: priority, rank -> chunk, envs (<= 32)
    __shared__ uint16_t cstart[136];                        // per chunk: first position in perm[]
    __shared__ int n_chunks;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 8) sh_stats[tid] = 0;
    if (tid < NBUCKET) hist[tid] = 0;
    if (tid == 0) { next_chunk = 0; n_reset = 0; }
    phase_stamp(B, 0, &next_chunk);
    stage_levels(levels, B.levels, B.n_levels, &bar);      // contains a __syncthreads()
    phase_stamp(B, 1, &next_chunk);

    const int64_t base = B.r_begin + (int64_t)blockIdx.x * tile;
    const int count = (int)min((int64_t)tile, B.r_begin + B.r_count - base);
    const uint32_t t_now = B.step_counter[0];               // gym steps taken before this call
    const uint32_t max_steps = B.max_steps > 0 ? (uint32_t)B.max_steps : 0xFFFFFFFFu;
    const int od = B.obs_dim;

    uint32_t st_cnt = 0;                                    // errors | episodes << 8 | successes << 16 | ran << 24
    int st_ticks = 0, st_ret = 0, st_epsteps = 0;

    // ---- phase A ---------------------------------------------------------------
    {
        // 16-byte accesses when this tile's slices start on 16 bytes (base is a multiple of 4 for every tile the
        // launchers produce; caller-provided arrays may be offset)
        const bool va = ((reinterpret_cast<uintptr_t>(actions + base) | reinterpret_cast<uintptr_t>(B.plan + base) |
                          reinterpret_cast<uintptr_t>(B.ep_start + base)) & 15u) == 0;
        const bool vo = (!reward || (reinterpret_cast<uintptr_t>(reward + base) & 15u) == 0) &&
                        (!done_out || (reinterpret_cast<uintptr_t>(done_out + base) & 3u) == 0) &&
                        (!ran_out || (reinterpret_cast<uintptr_t>(ran_out + base) & 3u) == 0) &&
                        (!avail_out || (reinterpret_cast<uintptr_t>(avail_out + base) & 7u) == 0);
        const uint32_t hist_s = (uint32_t)__cvta_generic_to_shared(hist), rank_s = (uint32_t)__cvta_generic_to_shared(rank);
        for (int el0 = 4 * tid; el0 < count; el0 += 4 * STEP_THREADS) {
            const int m = min(4, count - el0);
            const int64_t i0 = base + el0;
            int av[4]; uint32_t plo[4], phi[4], ev[4];
            if (va && m == 4) {
                const int4 a4 = *reinterpret_cast<const int4 *>(actions + i0);
                const uint4 p01 = *reinterpret_cast<const uint4 *>(B.plan + i0), p23 = *reinterpret_cast<const uint4 *>(B.plan + i0 + 2);
                const uint4 e4 = *reinterpret_cast<const uint4 *>(B.ep_start + i0);
                av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
                plo[0] = p01.x; phi[0] = p01.y; plo[1] = p01.z; phi[1] = p01.w; plo[2] = p23.x; phi[2] = p23.y; plo[3] = p23.z; phi[3] = p23.w;
                ev[0] = e4.x; ev[1] = e4.y; ev[2] = e4.z; ev[3] = e4.w;
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint64_t p = (u < m) ? B.plan[i0 + u] : 0ull;
                    av[u] = (u < m) ? actions[i0 + u] : TG_NUM_OPTIONS; plo[u] = (uint32_t)p; phi[u] = (uint32_t)(p >> 32);
                    ev[u] = (u < m) ? B.ep_start[i0 + u] : t_now + 1u;
                }
            }
            uint32_t dn = 0, cd = 0;                                   // 4 done bytes, 4 code bytes
            uint32_t special = 0;                                      // per element: bit u = reset, bit 4 + u = reference would raise
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t a = (uint32_t)av[u] < (uint32_t)TG_NUM_OPTIONS ? (uint32_t)av[u] : 12u;   // tg:92 raises IndexError; here: not run (bit 12 is never set)
                const uint32_t lo = plo[u];
                const bool run = (lo >> a) & 1u;
                // sort key: code-path class major (lanes of a warp then run the same policy and the same branch of
                // tick()), estimated length minor (longest first).  class of option a: nibble a of 0x332241100 (walk, ladder, drop, jump, interact)
                const uint32_t lenc = ((a < 4u ? lo >> (PL_LEN + 4u * a) : phi[u] >> (4u * a - 16u)) & 15u);
                const uint32_t bucket = (uint32_t)((0x332241100ull >> (4u * a)) & 15ull) * 12u + 11u - lenc;
                const uint32_t steps = t_now + 1u - ev[u];
                // branch-free: a warp's lanes disagree on `run` in nearly every element (20 % run), and a branch here costs both of
                // its sides plus the reconvergence for all of them; the rank is a predicated shared-memory atomic
                const uint32_t d = (run ? 0u : 0xFFu) & (((lo >> PL_TERM) & 1u) | (steps >= max_steps ? (uint32_t)TG_DONE_TRUNCATED : 0u));
                dn |= d << (8 * u);
                cd |= (run ? bucket : 255u) << (8 * u);
                asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 r;\n\tsetp.ne.u32 p, %0, 0;\n\t@p atom.shared.add.u32 r, [%1], 1;\n\t@p st.shared.u16 [%2], r;\n\t}"
                             :: "r"((uint32_t)run), "r"(hist_s + 4u * bucket), "r"(rank_s + 2u * (uint32_t)(el0 + u)) : "memory");
                special |= (d ? 1u : 0u) << u;
                // "the reference would raise": plan bit PL_ERR_DL + (a - TG_DOWN_LEFT), for the two drop options only (the plan never
                // sets it together with the option's can_run bit, compute_plan)
                static_assert(PL_ERR_DR == PL_ERR_DL + 1 && TG_DOWN_RIGHT == TG_DOWN_LEFT + 1, "error bits follow the option ids");
                special |= ((lo >> (a + (uint32_t)(PL_ERR_DL - TG_DOWN_LEFT))) & (((3u << TG_DOWN_LEFT) >> a) & 1u)) << (4 + u);
            }
            *reinterpret_cast<uint32_t *>(code + el0) = cd;            // el0 is a multiple of 4, cap a multiple of 16
            if (special) {                                             // rare: episode over (time limit / done), or the reference would raise
                for (int u = 0; u < m; u++) {
                    const int64_t i = i0 + u;
                    if ((special >> (4 + u)) & 1u) {
                        uint32_t *fw = reinterpret_cast<uint32_t *>(B.core + i) + 1;   // flag the env (target None in the reference)
                        const uint32_t f = *fw;
                        if (!(f & (1u << F_ERROR))) { *fw = f | (1u << F_ERROR); st_cnt += 1u; }
                    }
                    if ((special >> u) & 1u) {
                        if (B.auto_reset) rlist[atomicAdd(&n_reset, 1)] = (uint16_t)(el0 + u);      // statistics: when it is reset
                        else {
                            st_cnt += (1u << 8) + (((dn >> (8 * u)) & TG_DONE_TERMINATED) ? 1u << 16 : 0u);
                            st_ret += (int)B.acct[i].y; st_epsteps += (int)(t_now + 1u - ev[u]);
                        }
                    }
                }
            }
            if (B.flag_done) {                                         // bound flag views (cudaMalloc'ed by the caller's framework: 4-byte aligned rows of 4)
                if (m == 4 && ((base & 3) == 0)) {
                    *reinterpret_cast<uint32_t *>(B.flag_done + i0) = (dn | (dn >> 1)) & 0x01010101u;
                    *reinterpret_cast<uint32_t *>(B.flag_term + i0) = dn & 0x01010101u;
                    *reinterpret_cast<uint32_t *>(B.flag_trunc + i0) = (dn >> 1) & 0x01010101u;
                } else {
                    for (int u = 0; u < m; u++) {
                        const uint32_t d1 = (dn >> (8 * u)) & 255u;
                        B.flag_done[i0 + u] = d1 ? 1 : 0; B.flag_term[i0 + u] = d1 & 1u; B.flag_trunc[i0 + u] = (d1 >> 1) & 1u;
                    }
                }
            }
            if (vo && m == 4) {
                if (reward) *reinterpret_cast<float4 *>(reward + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (done_out) *reinterpret_cast<uint32_t *>(done_out + i0) = dn;
                if (ran_out) *reinterpret_cast<uint32_t *>(ran_out + i0) = 0u;
                if (avail_out) *reinterpret_cast<ushort4 *>(avail_out + i0) = make_ushort4(plo[0] & 0x1FFu, plo[1] & 0x1FFu, plo[2] & 0x1FFu, plo[3] & 0x1FFu);
            } else {
                for (int u = 0; u < m; u++) {
                    if (reward) reward[i0 + u] = 0.f;
                    if (done_out) done_out[i0 + u] = (uint8_t)(dn >> (8 * u));
                    if (ran_out) ran_out[i0 + u] = 0;
                    if (avail_out) avail_out[i0 + u] = (uint16_t)(plo[u] & 0x1FFu);
                }
            }
        }
    }
    __syncthreads();
    phase_stamp(B, 2, &next_chunk);
    // ---- sort: exclusive scan of the 64 bucket counts (warp 0), scatter the runnable envs by rank, then order their
    // 32-env chunks longest first across classes (a chunk's first env is its longest: length is the minor key) ----
    if (tid < 32) scan64(hist, lane);
    __syncthreads();
    for (int el = tid; el < count; el += STEP_THREADS) {
        const int c = code[el];
        if (c != 255) {
            perm[hist[c] + rank[el]] = (uint16_t)el;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(B.core + base + el));     // phase B loads these two records (one
            asm volatile("prefetch.global.L2 [%0];" :: "l"(B.acct + base + el));     // sector each) at the head of its chunk
        }
    }
    const int n_run = hist[NBUCKET - 1];                    // exclusive scan: start of the (unused) last bucket = runnable envs
    const int n_idle_rst = n_reset;                         // idle envs to reset; the option lanes append theirs
    __syncthreads();
    if (tid == STEP_THREADS - 1 && B.sp_count) {            // sparse outputs: this tile's block of records, and where it is
        const uint32_t nrec = (uint32_t)(n_run + n_idle_rst);
        const uint32_t rb = nrec ? atomicAdd(B.sp_count, nrec) : 0u;
        rec_base = rb;
        *reinterpret_cast<uint2 *>(B.sp_count + 4 + 2 * blockIdx.x) = make_uint2(rb, nrec);   // tile table: after the 16-byte header
    }
    // Chunks of up to 32 envs that never straddle two classes (a warp whose lanes belong to different classes runs their
    // code paths one after the other: the chunk at the ladder / drop / jump / interact border took 100 us): per class,
    // ceil(count / 32) chunks.  Order: drops, jumps and interact first (the longest serial chains, and the interact tick's cold code path), then
    // longest first across classes (a chunk's first env is its longest: length is the minor sort key).
    if (tid < 32) {
        const int cs = lane < 5 ? hist[12 * lane] : n_run, ce = lane < 5 ? hist[12 * (lane + 1)] : n_run;   // class lane = perm[cs .. ce)
        const int nch = (ce - cs + 31) >> 5;
        int incl = nch;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
        const int q0 = incl - nch;
        const int nq = __shfl_sync(0xFFFFFFFFu, incl, 4);
        for (int t = 0; t < nch; t++) {
            cstart[q0 + t] = (uint16_t)(cs + 32 * t);
            clen[q0 + t] = (uint8_t)min(32, ce - cs - 32 * t);
            ckey[q0 + t] = (uint8_t)(11 - code[perm[cs + 32 * t]] % 12 + ((lane >= 2 && lane <= 4) ? 12 : 0));
        }
        if (lane == 0) n_chunks = nq;
        __syncwarp();
        for (int c = lane; c < nq; c += 32) {
            const int k = ckey[c];
            int r = 0;
            for (int j = 0; j < nq; j++) { const int kj = ckey[j]; r += (kj > k || (kj == k && j < c)) ? 1 : 0; }
            order[r] = (uint8_t)c;
        }
    }
    __syncthreads();
    const int nrc = n_chunks;
    phase_stamp(B, 3, &next_chunk);

    // ---- phase B: 32-env chunks of runnable envs (class-sorted, longest first) from a shared counter
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(&next_chunk, 1);
        q = __shfl_sync(0xFFFFFFFFu, q, 0);
        if (q >= nrc) break;
        const int cq = order[q];
        const int j = cstart[cq] + lane;
        unsigned long long chunk_t0 = 0;                                          // debug instrumentation: per-chunk timing
        if (B.phase_ts && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(chunk_t0) :: "memory");
        const bool valid = lane < clen[cq];
        const int el = valid ? perm[j] : 0;
        const int64_t i = base + el;
        const int lid = (valid && B.level_id) ? B.level_id[i] : 0;
        const LevelBlob &L = levels[lid];
        const int a = valid ? actions[i] : -1;
        const uint32_t eps = valid ? B.ep_start[i] : 0u;
        Env<NI> e;
        uint4 acct;
        int n, tcx = 0;
        uint32_t newerr = 0, err0 = 0, drawn_i = 0;
        if (a == TG_INTERACT) drawn_i = interact_option_mem<TAPE, NI>(B, &L, i);   // whole chunks: interact is its own sort class
        if (valid) {
            load_env(e, B, i, acct);
            if (a == TG_INTERACT) e.d0 = e.draws - drawn_i;
            else { bool err; option_setup(e, L, a, tcx, err); err0 = e.flags; }   // runnable (the plan says so): target column; flags before the option
        }
        unsigned long long chunk_t1 = 0, chunk_t2 = 0, chunk_t3 = 0;
        if (B.phase_ts && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(chunk_t1) : "r"(tcx + (int)acct.x) : "memory");
        n = run_option_to_end<TAPE, NI, false>(e, L, a, tcx, valid && a != TG_INTERACT);
        if (B.phase_ts && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(chunk_t2) : "r"(n) : "memory");
        if (!valid) continue;
        if (a == TG_INTERACT) n = 1;
        else newerr = (e.flags & ~err0 & (1u << F_ERROR)) ? 1u : 0u;
        st_cnt += newerr + (1u << 24);
        st_ticks += n;
        const bool jump = a >= TG_JUMP_LEFT;
        const int r = -n - (jump ? 4 : 0);                                        // impl:15-16: -1 per tick, JUMP tick -5
        acct.y = (uint32_t)((int)acct.y + r);
        const uint32_t steps = t_now + 1u - eps;
        const bool term = is_done(e, L);
        const uint32_t d = (term ? TG_DONE_TERMINATED : 0) | (steps >= max_steps ? TG_DONE_TRUNCATED : 0);
        if (d) { st_cnt += (1u << 8) + (term ? 1u << 16 : 0u); st_ret += (int)acct.y; st_epsteps += (int)steps; }
        if (d && B.auto_reset) {                                                  // reset in phase C (full warps)
            acct.y = 0; acct.z = pack_pending(e.draws - e.d0, n, jump, d);
            store_env(e, B, i, acct);
            rank[el] = (uint16_t)j;
            rlist[atomicAdd(&n_reset, 1)] = (uint16_t)(el | 0x8000);
        } else {
            store_env(e, B, i, acct);
            const uint64_t plan = plan_of(e, L);
            B.plan[i] = plan;
            if (B.phase_ts && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(chunk_t3) : "r"((uint32_t)plan) : "memory");
            uint32_t *rec = B.sp_count ? B.sp_recs + (size_t)(rec_base + (uint32_t)j) * B.sp_words : nullptr;
            if (rec) { rec[0] = (uint32_t)i; rec[1] = __float_as_uint((float)r); rec[2] = d | 256u; }
            if (obs || rec)
                // An option other than interact that picked nothing up moved the player and nothing else: two slots of the
                // row (it holds the env's previous observation, tg_step in treasure_b200.h).  Records carry whole rows.
                write_obs(e, L, B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, obs ? obs + i * od : nullptr, od,
                          rec ? reinterpret_cast<float *>(rec + 3) : nullptr,
                          rec || a == TG_INTERACT || (((e.flags ^ err0) >> F_ERROR) | ((e.flags >> F_ERROR) & 1u)) != 0u);
            if (avail_out) avail_out[i] = (uint16_t)(plan & 0x1FFu);
        }
        if (reward) reward[i] = (float)r;
        if (done_out) done_out[i] = (uint8_t)d;
        if (ran_out) ran_out[i] = (uint8_t)1;
        if (B.flag_done) { B.flag_done[i] = d ? 1 : 0; B.flag_term[i] = d & 1u; B.flag_trunc[i] = (d >> 1) & 1u; }
        if (B.phase_ts && lane == 0 && q < 128 && blockIdx.x < 8192) {                                 // [8192][8] phase stamps, then [grid][128][4] chunks
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1) :: "memory");
            unsigned long long *ct = B.phase_ts + (size_t)8192 * 8 + ((size_t)blockIdx.x * 128 + (size_t)q) * 4;
            ct[0] = chunk_t0; ct[1] = (t1 - chunk_t0) | ((unsigned long long)(code[el] / 12) << 32) | ((unsigned long long)n << 40);
            ct[2] = (chunk_t1 - chunk_t0) | ((chunk_t2 - chunk_t1) << 32);
            ct[3] = chunk_t3 ? ((chunk_t3 - chunk_t2) | ((t1 - chunk_t3) << 32)) : 0ull;
        }
    }
    phase_stamp(B, 4, &next_chunk);
    __syncthreads();
    phase_stamp(B, 5, &next_chunk);

    // ---- phase C: the envs whose episode ended (auto_reset): reset_game (impl:55-73) with full warps ----
    const int n_rst = n_reset;
    for (int k = tid; k < n_rst; k += STEP_THREADS) {
        const uint32_t v = rlist[k];
        const int el = (int)(v & 0x7FFFu);
        const int64_t i = base + el;
        const int lid = B.level_id ? B.level_id[i] : 0;
        const LevelBlob &L = levels[lid];
        Env<NI> e;
        uint4 acct;
        load_env(e, B, i, acct);
        uint32_t drawn = 0, rec_reward = 0, rec_flags, slot;
        if (v & 0x8000u) {                                                        // its option ran: statistics already counted
            const uint32_t z = acct.z;
            drawn = z & 0xFFFFu;
            const int r = -(int)((z >> 16) & 0x1FFFu) - (((z >> 29) & 1u) ? 4 : 0);
            rec_reward = __float_as_uint((float)r); rec_flags = (z >> 30) | 256u;
            slot = rank[el];
        } else {
            const uint32_t steps = t_now + 1u - B.ep_start[i];
            const bool term = (B.plan[i] >> PL_TERM) & 1ull;
            rec_flags = (term ? TG_DONE_TERMINATED : 0) | (steps >= max_steps ? TG_DONE_TRUNCATED : 0);
            st_cnt += (1u << 8) + (term ? 1u << 16 : 0u); st_ret += (int)acct.y; st_epsteps += (int)steps;
            slot = (uint32_t)(n_run + k);                                         // k < n_idle_rst: the list starts with the idle envs
        }
        acct.y = 0; acct.z = 0;
        // the Philox blocks of a call are numbered from the call's first draw (draw_w): the reset continues the numbering
        e.d0 = e.draws - drawn;
        if (!TAPE && (drawn & 3u)) { const uint4 o = philox_block(e.d0, drawn >> 2, e.id_lo, e.id_hi, e.key0, e.key1); e.w0 = o.x; e.w1 = o.y; e.w2 = o.z; e.w3 = o.w; }
        reset_env<TAPE>(e, L);
        store_env(e, B, i, acct);
        const uint64_t plan = plan_of(e, L);
        B.plan[i] = plan;
        B.ep_start[i] = t_now + 1u;
        uint32_t *rec = B.sp_count ? B.sp_recs + (size_t)(rec_base + slot) * B.sp_words : nullptr;
        if (rec) { rec[0] = (uint32_t)i; rec[1] = rec_reward; rec[2] = rec_flags; }
        if (obs || rec)
            write_obs(e, L, B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, obs ? obs + i * od : nullptr, od,
                      rec ? reinterpret_cast<float *>(rec + 3) : nullptr);
        if (avail_out) avail_out[i] = (uint16_t)(plan & 0x1FFu);
    }
    phase_stamp(B, 6, &next_chunk);
    int st[8];
    st[ST_EPISODES] = (st_cnt >> 8) & 255; st[ST_SUCCESS] = (st_cnt >> 16) & 255; st[ST_RETURN] = st_ret; st[ST_EPSTEPS] = st_epsteps;
    st[ST_TICKS] = st_ticks; st[ST_RAN] = st_cnt >> 24; st[ST_ERRORS] = st_cnt & 255;
    st[ST_STEPS] = (tid == 0) ? count : 0;                  // every env of the tile takes exactly one gym step
    stats_accumulate(sh_stats, B.stats, st);
    phase_stamp(B, 7, &next_chunk);
    finish_launch(B);
}

// Observation rows of every env (impl:368-378) from the stored state: follows a step kernel whose caller's `obs` is
// not known to hold the previous observations, and serves tg_reset.  Rows are built in shared memory and leave
// with full-line stores.
template <int NI>
__global__ void __launch_bounds__(256)
tg_obs_kernel(const __grid_constant__ BatchView B, float *__restrict__ obs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    float *stage = reinterpret_cast<float *>(smem_raw + (((size_t)B.n_levels * sizeof(LevelBlob) + 15) & ~(size_t)15));   // [256][obs_dim]
    __shared__ uint64_t bar;
    stage_levels(levels, B.levels, B.n_levels, &bar);
    const int od = B.obs_dim;
    const int64_t i0 = B.r_begin + (int64_t)blockIdx.x * 256, i = i0 + threadIdx.x;
    const int rows = (int)min((int64_t)256, B.r_begin + B.r_count - i0);
    if ((int)threadIdx.x < rows) {
        const int lid = B.level_id ? B.level_id[i] : 0;
        Env<NI> e;
        load_core(e, B, i);
        e.angles = B.angles + i; e.n = B.n;
        write_obs(e, levels[lid], B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, stage + threadIdx.x * od, od);
    }
    __syncthreads();
    float *dst = obs + i0 * od;
    const int nf = rows * od;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (nf & 3) == 0) {
        for (int k = threadIdx.x; k < (nf >> 2); k += 256) reinterpret_cast<float4 *>(dst)[k] = reinterpret_cast<const float4 *>(stage)[k];
    } else {
        for (int k = threadIdx.x; k < nf; k += 256) dst[k] = stage[k];
    }
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
        acct.y = 0;
        store_env(e, B, i, acct);
        B.plan[i] = plan_of(e, L);
        B.ep_start[i] = B.step_counter[0];                 // the next gym step is the first of the new episode
    }
    if (obs) write_obs(e, L, B.obs_lut + (size_t)(B.level_id ? B.level_id[i] : 0) * 2 * OBS_LUT_N, obs + i * B.obs_dim, B.obs_dim);
}

// TreasureGame.available_mask (tg:83-89): the nine can_run bits live in the env's plan word
__global__ void __launch_bounds__(256)
tg_mask_kernel(BatchView B, uint8_t *__restrict__ mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B.n) return;
    const uint32_t m = (uint32_t)(B.plan[i] & 0x1FFull);
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
    const uint32_t t_now = B.step_counter[0];
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
        const uint32_t steps = t_now + 1u - B.ep_start[i];
        const bool term = is_done(e, L);
        const bool trunc = B.max_steps > 0 && steps >= (uint32_t)B.max_steps;
        const int d = (term ? TG_DONE_TERMINATED : 0) | (trunc ? TG_DONE_TRUNCATED : 0);
        st[ST_TICKS] = 1; st[ST_RAN] = 1; st[ST_STEPS] = 1;
        st[ST_ERRORS] = ((e.flags & (1u << F_ERROR)) && !err0) ? 1 : 0;
        if (d) {
            st[ST_EPISODES] = 1; st[ST_SUCCESS] = term; st[ST_RETURN] = (int)acct.y; st[ST_EPSTEPS] = (int)steps;
            if (B.auto_reset) { reset_env<TAPE>(e, L); acct.y = 0; B.ep_start[i] = t_now + 1u; }
        }
        store_env(e, B, i, acct);
        B.plan[i] = plan_of(e, L);
        if (obs) write_obs(e, L, B.obs_lut + (size_t)(B.level_id ? B.level_id[i] : 0) * 2 * OBS_LUT_N, obs + i * B.obs_dim, B.obs_dim);
        if (reward) reward[i] = (float)r;
        if (done_out) done_out[i] = (uint8_t)d;
    }
    stats_accumulate(sh_stats, B.stats, st);
    finish_launch(B);
}

// tg_step_frames: the option layer with a drawer (opt:20-36 with opt:33-34: draw_domain() after every primitive tick).
// The M selected envs take one gym step; their option runs tick by tick in the general loop (no multi-tick shortcuts) and
// after tick t the env's state is copied to slot [k][t-1] of the snapshot arrays S (core / items / angles / level id, laid
// out like a batch of M * T envs), which the renderer then draws.  Slots after the option's last tick repeat the final
// state of the option (before any auto-reset).  Accounting as in tg_step_kernel; the batch-wide step counter does not move
// (only these envs stepped), so their episode start is moved back by one instead.
template <bool TAPE, int NI>
__global__ void __launch_bounds__(AUX_THREADS)
tg_trace_kernel(BatchView B, BatchView SN, const int64_t *__restrict__ env_ids, int M, const int32_t *__restrict__ actions, int T,
                int32_t *__restrict__ n_ticks, float *__restrict__ obs, float *__restrict__ reward, uint8_t *__restrict__ done_out,
                uint8_t *__restrict__ ran_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LevelBlob *levels = reinterpret_cast<LevelBlob *>(smem_raw);
    __shared__ uint64_t bar;
    stage_levels(levels, B.levels, B.n_levels, &bar);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    const int64_t i = env_ids[k];
    if (i < 0 || i >= B.n) { n_ticks[k] = -1; return; }
    const int lid = B.level_id ? B.level_id[i] : 0;
    const LevelBlob &L = levels[lid];
    const uint32_t t_now = B.step_counter[0];
    Env<NI> e;
    uint4 acct;
    load_env(e, B, i, acct);
    auto snap = [&](int slot) {
        const int64_t q = (int64_t)k * T + slot;
        uint4 c;
        c.x = pack_player(e.px, e.py, e.sticky); c.y = e.flags; c.z = pack_xy(e.ix[0], e.iy[0]);
        c.w = (NI > 1) ? pack_xy(e.ix[NI > 1 ? 1 : 0], e.iy[NI > 1 ? 1 : 0]) : 0u;
        SN.core[q] = c;
        if (NI > 2 && SN.items23) SN.items23[q] = make_uint2(pack_xy(e.ix[NI > 2 ? 2 : 0], e.iy[NI > 2 ? 2 : 0]), NI > 3 ? pack_xy(e.ix[NI > 3 ? 3 : 0], e.iy[NI > 3 ? 3 : 0]) : 0u);
        for (int h = 0; h < TG_MAX_HANDLES; h++) SN.angles[(int64_t)h * SN.n + q] = B.angles[(int64_t)h * B.n + i];
        if (SN.level_id) const_cast<uint8_t *>(SN.level_id)[q] = (uint8_t)lid;
    };
    const int a = actions[k];
    int tcx = 0; bool err = false;
    const bool runnable = (unsigned)a < (unsigned)TG_NUM_OPTIONS && option_setup(e, L, a, tcx, err);
    const uint32_t err0 = e.flags & (1u << F_ERROR);
    if (err) e.flags |= 1u << F_ERROR;
    int n = 0;
    if (runnable) {
        const int s = (a == TG_GO_LEFT || a == TG_DOWN_LEFT || a == TG_JUMP_LEFT) ? -1 : 1;
        general_option_loop<TAPE, NI, true, false>(e, L, a, s, tcx * S + S / 2, n, [&](int t) { if (t <= T) snap(t - 1); });
    }
    for (int t = min(n, T); t < T; t++) snap(t);
    n_ticks[k] = n;
    const int r = runnable ? -n - ((a >= TG_JUMP_LEFT) ? 4 : 0) : 0;
    acct.y = (uint32_t)((int)acct.y + r);
    const uint32_t steps = t_now + 1u - B.ep_start[i];
    const bool term = is_done(e, L), trunc = B.max_steps > 0 && steps >= (uint32_t)B.max_steps;
    const uint32_t d = (term ? TG_DONE_TERMINATED : 0) | (trunc ? TG_DONE_TRUNCATED : 0);
    atomicAdd(&B.stats[ST_STEPS], 1ull);
    if (runnable) { atomicAdd(&B.stats[ST_RAN], 1ull); atomicAdd(&B.stats[ST_TICKS], (unsigned long long)n); }
    if ((e.flags & (1u << F_ERROR)) && !err0) atomicAdd(&B.stats[ST_ERRORS], 1ull);
    uint32_t ep0 = B.ep_start[i] - 1u;                                           // one more gym step than the batch counter says
    if (d) {
        atomicAdd(&B.stats[ST_EPISODES], 1ull);
        if (term) atomicAdd(&B.stats[ST_SUCCESS], 1ull);
        atomicAdd(&B.stats[ST_RETURN], (unsigned long long)(long long)(int)acct.y);
        atomicAdd(&B.stats[ST_EPSTEPS], (unsigned long long)steps);
        if (B.auto_reset) { reset_env<TAPE>(e, L); acct.y = 0; ep0 = t_now; }
    }
    store_env(e, B, i, acct);
    B.plan[i] = plan_of(e, L);
    B.ep_start[i] = ep0;
    if (obs) write_obs(e, L, B.obs_lut + (size_t)lid * 2 * OBS_LUT_N, obs + (int64_t)k * B.obs_dim, B.obs_dim);
    if (reward) reward[k] = (float)r;
    if (done_out) done_out[k] = (uint8_t)d;
    if (ran_out) ran_out[k] = runnable ? 1 : 0;
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
    B.plan[i] = plan_of(e, L);
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
    if (v.acct) {
        v.acct[i * 3] = (int)a.y; v.acct[i * 3 + 1] = B.step_counter[0] - B.ep_start[i];
        v.acct[i * 3 + 2] = ((f >> F_ERROR) & 1u) | (core_sticky(c.x) << 1);
    }
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
        a.y = (uint32_t)(int)v.acct[i * 3];
        B.ep_start[i] = B.step_counter[0] - (uint32_t)v.acct[i * 3 + 1];
        f = (f & ~(1u << F_ERROR)) | (((v.acct[i * 3 + 2] & 1) ? 1u : 0u) << F_ERROR);
        c.x = pack_player(core_px(c.x), hi16(c.x), (uint32_t)(v.acct[i * 3 + 2] >> 1));
    }
    c.y = f;
    B.core[i] = c; B.acct[i] = a;
    uint2 h = make_uint2(0u, 0u);
    if (B.items23) h = B.items23[i];
    B.plan[i] = compute_plan<4>(&B.levels[B.level_id ? B.level_id[i] : 0], pack_player(core_px(c.x), hi16(c.x), 0u), f, c.z, c.w, h.x, h.y);
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }
static inline size_t level_smem(const BatchView &B) { return (size_t)B.n_levels * sizeof(LevelBlob); }
// step kernel: level blobs + the tile's sort arrays (code u8, rank / perm / reset list u16)
static inline size_t step_smem(const BatchView &B, int cap) { return ((level_smem(B) + 15) & ~(size_t)15) + (size_t)cap * 7; }

template <bool TAPE, int NI>
static cudaError_t step_impl(const BatchView &B, int tile, const int32_t *a, float *obs, float *rew, uint8_t *done,
                             uint8_t *ran, uint16_t *avail, cudaStream_t s) {
    const int cap = (tile + 15) & ~15;
    const size_t smem = step_smem(B, cap);
    if (smem > 40 * 1024) {   // static + dynamic shared memory can pass the default 48 KB (many layouts): opt in, once per size and device
        static size_t allowed[MAX_DEVICES] = {};
        const int dslot = device_slot();
        if (smem > allowed[dslot]) {
            cudaError_t r = cudaFuncSetAttribute(tg_step_kernel<TAPE, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (r != cudaSuccess) return r;
            allowed[dslot] = smem;
        }
    }
    tg_step_kernel<TAPE, NI><<<grid_for(B.r_count, tile), STEP_THREADS, smem, s>>>(B, tile, cap, a, obs, rew, done, ran, avail);
    return cudaGetLastError();
}

// Tile size: large tiles sort better (more runnable envs per tile -> fuller warps); small tiles give more
// CTAs.  The grid is sized in whole "slots": 148 SMs x 3 resident CTAs = 444 CTAs run at once, so the tile is
// n / (444 * waves) rounded up (2364 envs for a 1,048,576-env step) -- with fixed 2048-env tiles and 4 CTAs per SM a
// step had 512 CTAs, 68 SMs held four of them and 80 SMs three.  Smaller batches aim for two, then one CTA per SM slot.  TG_STEP_TILE overrides.
int pick_step_tile(int64_t n, int forced) {
    if (forced >= 32 && forced <= 4096) return (forced + 3) / 4 * 4;
    const int slots = device_sm_count() * TG_STEP_MIN_BLOCKS;
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

cudaError_t launch_step(const BatchView &B, int ni, int forced_tile, const int32_t *a, float *obs, float *rew, uint8_t *done,
                        uint8_t *ran, uint16_t *avail, cudaStream_t s) {
    const bool tape = B.tape != nullptr;
    const int tile = pick_step_tile(B.r_count, forced_tile);
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

cudaError_t launch_trace(const BatchView &B, const BatchView &S, int ni, const int64_t *env_ids, int M, const int32_t *actions, int T,
                         int32_t *n_ticks, float *obs, float *rew, uint8_t *done, uint8_t *ran, cudaStream_t s) {
    const unsigned g = grid_for(M, AUX_THREADS);
    const size_t sm = level_smem(B);
    const bool tape = B.tape != nullptr;
    if (ni <= 2) {
        if (tape) tg_trace_kernel<true, 2><<<g, AUX_THREADS, sm, s>>>(B, S, env_ids, M, actions, T, n_ticks, obs, rew, done, ran);
        else tg_trace_kernel<false, 2><<<g, AUX_THREADS, sm, s>>>(B, S, env_ids, M, actions, T, n_ticks, obs, rew, done, ran);
    } else {
        if (tape) tg_trace_kernel<true, 4><<<g, AUX_THREADS, sm, s>>>(B, S, env_ids, M, actions, T, n_ticks, obs, rew, done, ran);
        else tg_trace_kernel<false, 4><<<g, AUX_THREADS, sm, s>>>(B, S, env_ids, M, actions, T, n_ticks, obs, rew, done, ran);
    }
    return cudaGetLastError();
}

cudaError_t launch_mask(const BatchView &B, int ni, uint8_t *mask, cudaStream_t s) {
    tg_mask_kernel<<<grid_for(B.n, 256), 256, 0, s>>>(B, mask);
    return cudaGetLastError();
}

cudaError_t launch_obs(const BatchView &B, int ni, float *obs, cudaStream_t s) {
    const size_t smem = ((level_smem(B) + 15) & ~(size_t)15) + (size_t)256 * B.obs_dim * sizeof(float);
    if (smem > 40 * 1024) {
        static size_t allowed[MAX_DEVICES][2] = {};
        const int dslot = device_slot(), v = ni <= 2 ? 0 : 1;
        if (smem > allowed[dslot][v]) {
            cudaError_t r = ni <= 2 ? cudaFuncSetAttribute(tg_obs_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                    : cudaFuncSetAttribute(tg_obs_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (r != cudaSuccess) return r;
            allowed[dslot][v] = smem;
        }
    }
    if (ni <= 2) tg_obs_kernel<2><<<grid_for(B.r_count, 256), 256, smem, s>>>(B, obs);
    else tg_obs_kernel<4><<<grid_for(B.r_count, 256), 256, smem, s>>>(B, obs);
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
