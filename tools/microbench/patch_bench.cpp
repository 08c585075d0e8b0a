// CPU-only timing of the host half of tg_step_host_sparse (csrc/tg_host_patch.h): synthetic records with the step
// kernel's layout (per tile a contiguous block, runnable envs in arbitrary order), 20 % of the envs reported per step.
//   g++ -O3 -fopenmp -march=native -I gym_treasure_game_b200/csrc tools/microbench/patch_bench.cpp -o /tmp/patch_bench
#include <cstdio>
#include <random>
#include <vector>
#include "tg_host_patch.h"
int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 1 << 20;
    const int tile = 3544, od = 9, words = 12, steps = 30;
    const int grid = (int)((n + tile - 1) / tile);
    std::vector<float> obs((size_t)n * od), reward(n);
    std::vector<uint8_t> done(n), ran(n);
    std::mt19937 rng(1);
    std::vector<std::vector<uint32_t>> recs(steps), table(steps);
    for (int s = 0; s < steps; s++) {
        recs[s].reserve((size_t)n / 4 * words); table[s].resize(2 * (size_t)grid);
        uint32_t nr = 0;
        for (int t = 0; t < grid; t++) {
            const int64_t base = (int64_t)t * tile; const int m = (int)std::min<int64_t>(tile, n - base);
            std::vector<uint32_t> els;
            for (int el = 0; el < m; el++) if (rng() % 5 == 0) els.push_back(el);
            std::shuffle(els.begin(), els.end(), rng);
            table[s][2 * t] = nr; table[s][2 * t + 1] = (uint32_t)els.size();
            for (uint32_t el : els) { uint32_t r[12] = {(uint32_t)(base + el), 0xC0000000u, 256u}; for (int k = 3; k < 12; k++) r[k] = rng(); recs[s].insert(recs[s].end(), r, r + 12); nr++; }
        }
    }
    double best = 1e9, sum = 0;
    for (int rep = 0; rep < 3; rep++)
        for (int s = 0; s < steps; s++) {
            const double t0 = now_s();
            const bool ok = sparse_apply_tiles(table[s].data(), grid, tile, 0, n, recs[s].data(), words, od, obs.data(), reward.data(), done.data(), ran.data());
            const double dt = now_s() - t0;
            if (!ok) { printf("bad\n"); return 1; }
            if (rep) { sum += dt; best = std::min(best, dt); }
        }
    size_t nran = 0; for (int64_t i = 0; i < n; i++) nran += ran[i];
    printf("n=%lld threads=%d: patch mean %.3f ms  best %.3f ms  (%zu records per step; ran set on %zu envs)\n", (long long)n, host_threads(),
           1e3 * sum / (2 * steps), 1e3 * best, recs[0].size() / words, nran);
    return 0;
}
