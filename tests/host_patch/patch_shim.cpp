// Test shim: the host half of tg_step_host_sparse (csrc/tg_host_patch.h) behind a C entry point, compiled with g++ by
// tests/test_host_patch.py (no CUDA needed).
#include "tg_host_patch.h"
extern "C" int patch_apply(const uint32_t *table, int grid, int tile, long long lo, long long cnt, const uint32_t *recs, int words,
                           int od, float *obs, float *reward, uint8_t *done, uint8_t *ran) {
    return sparse_apply_tiles(table, grid, tile, lo, cnt, recs, words, od, obs, reward, done, ran) ? 0 : 1;
}
