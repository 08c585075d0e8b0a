// Host side of tg_step_host_sparse: patching the caller's arrays from the records of a chunk's launch.  Plain C++ +
// OpenMP (no CUDA types), so that tools/microbench/patch_bench.cpp can time it without a GPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <omp.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// Threads for the host-side patching: the host cores divided among the ranks that share the host (LOCAL_WORLD_SIZE of
// a torchrun launch; a process launched on its own takes all of them), minus one when the host is shared, 2..16; TG_HOST_THREADS overrides.  Passed as a num_threads clause because
// launchers such as torchrun export OMP_NUM_THREADS=1, which would leave one thread to scatter 200 k rows.
__attribute__((used)) static int host_threads() {   // referenced from OpenMP clauses only (the CUDA front end does not see those)
    static int t = 0;
    if (!t) {
        const char *v = getenv("TG_HOST_THREADS");
        int want = v ? atoi(v) : 0;
        if (want < 1) {
            int share = 0;
            const char *lw = getenv("LOCAL_WORLD_SIZE");
            if (lw) share = atoi(lw);
            if (share < 1) share = 1;                      // a single process: the host is ours (not: one share per visible device)
            want = omp_get_num_procs() / share;
            // several ranks on one host: leave a core per rank to everything that is not patching (the Python thread of the
            // next part, CUDA's helper threads).  Measured with 8 ranks on 32 cores, 131,072 envs each: 3 threads per rank
            // patch in 0.13 ms per step, 4 threads in 0.1 to 15 ms (spinning OpenMP teams oversubscribe the host).
            if (share > 1 && want > 2) want -= 1;
            if (want > 16) want = 16;
            if (want < 2) want = 2;
        }
        t = want > 64 ? 64 : want;
    }
    return t;
}

static inline double now_s() { return omp_get_wtime(); }

// bit u of the result = byte u of ran[] | done[] is non-zero, for 64 envs starting at i (w = how many are valid)
static inline uint64_t nonzero_mask64(const uint8_t *ran, const uint8_t *done, int w) {
    uint64_t m = 0;
#if defined(__SSE2__)
    if (w == 64) {
        const __m128i z = _mm_setzero_si128();
        for (int q = 0; q < 4; q++) {
            const __m128i x = _mm_or_si128(_mm_loadu_si128((const __m128i *)(ran + 16 * q)), _mm_loadu_si128((const __m128i *)(done + 16 * q)));
            m |= (uint64_t)((~_mm_movemask_epi8(_mm_cmpeq_epi8(x, z))) & 0xFFFF) << (16 * q);
        }
        return m;
    }
#endif
    for (int u = 0; u < w; u++) m |= (uint64_t)((ran[u] | done[u]) != 0) << u;
    return m;
}

// Host side of the sparse step, tile by tile (OpenMP over the tiles of a chunk's launch).  A tile's records are
// contiguous but in the kernel's sorted order; applied in that order every record would touch four scattered cache lines
// of the caller's arrays.  Instead: (a) scatter the record numbers into a tile-sized table (L1), (b) sweep the tile's envs
// in index order: a reported env takes its record; an env that is not reported but still shows the previous step's
// outputs (ran or done set) goes back to reward 0 / done 0 / ran 0 -- its observation row is already right, an option
// that cannot run leaves its env untouched (_option.py:22-23).  The caller's arrays are read and written front to back.
// Returns false when a record names an env outside its tile.
static bool sparse_apply_tiles(const uint32_t *table, int grid, int tile, int64_t lo, int64_t cnt, const uint32_t *recs, int words,
                               int od, float *obs, float *reward, uint8_t *done, uint8_t *ran) {
    bool ok = true;
#pragma omp parallel num_threads(host_threads()) if (cnt > 8192)
    {
        uint16_t pos[4096];                         // record number + 1 of the tile's env el (only read where `present` says so)
        uint64_t present[4096 / 64];
#pragma omp for schedule(dynamic, 2)
        for (int t = 0; t < grid; t++) {
            const int64_t base = lo + (int64_t)t * tile;
            const int m = (int)std::min<int64_t>(tile, lo + cnt - base);
            const uint32_t *tr = recs + (size_t)table[2 * t] * words;
            const uint32_t nrec = table[2 * t + 1];
            if (nrec > (uint32_t)m || m > 4096) { ok = false; continue; }
            memset(present, 0, sizeof(uint64_t) * (size_t)((m + 63) / 64));
            for (uint32_t r = 0; r < nrec; r++) {
                const uint32_t el = tr[(size_t)r * words] - (uint32_t)base;
                if (el >= (uint32_t)m) { ok = false; continue; }
                pos[el] = (uint16_t)r;
                // the rows this record will overwrite: on their way while the rest of the tile is scanned (the sweep below is
                // bound by how many cache lines a core keeps in flight)
                const float *o = obs + (size_t)(base + el) * od;
                __builtin_prefetch(o, 1); __builtin_prefetch(o + od - 1, 1); __builtin_prefetch(&reward[base + el], 1);
                present[el >> 6] |= 1ull << (el & 63);
            }
            for (int el0 = 0; el0 < m; el0 += 64) {
                const int64_t i0 = base + el0;
                const uint64_t P = present[el0 >> 6];
                uint64_t S = nonzero_mask64(ran + i0, done + i0, std::min(64, m - el0)) & ~P;   // shows the previous step, not reported now
                for (uint64_t b = P; b; b &= b - 1) {
                    const int u = __builtin_ctzll(b);
                    const int64_t i = i0 + u;
                    const uint32_t *rec = tr + (size_t)pos[el0 + u] * words;
                    memcpy(&reward[i], &rec[1], 4);
                    done[i] = (uint8_t)(rec[2] & 255u);
                    ran[i] = (uint8_t)(rec[2] >> 8);
                    if (od == 9) memcpy(obs + (size_t)i * 9, rec + 3, 36);      // the reference level: fixed-size copy, inlined
                    else memcpy(obs + (size_t)i * od, rec + 3, (size_t)od * 4);
                }
                for (; S; S &= S - 1) {
                    const int64_t i = i0 + __builtin_ctzll(S);
                    reward[i] = 0.0f; done[i] = 0; ran[i] = 0;
                }
            }
        }
    }
    return ok;
}

