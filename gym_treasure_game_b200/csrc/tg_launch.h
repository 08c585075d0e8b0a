// Kernel launchers shared between the translation units of libtreasure_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "tg_types.h"

namespace tg {

// Function attributes (dynamic shared memory opt-in) and SM counts are per device, and one process may hold envs on
// several devices: everything the launchers cache is indexed by the current device ordinal.
constexpr int MAX_DEVICES = 64;
inline int device_slot() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= MAX_DEVICES) d = 0;
    return d;
}
inline int device_sm_count() {
    static int sms[MAX_DEVICES] = {};
    const int d = device_slot();
    if (!sms[d]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v < 1) { cudaGetLastError(); v = 148; }
        sms[d] = v;
    }
    return sms[d];
}

// render assets of one level on the device
struct RenderAssets {
    const uint8_t *background;   // [frame_h][frame_w][3]
    const uint32_t *sprites;     // [TG_NUM_SPRITES][48][48] packed R | G<<8 | B<<16 | A<<24
};

struct RenderView {
    RenderAssets assets[TG_MAX_LEVELS];
    int32_t frame_w, frame_h, cw, ch;
    unsigned long long *job_counter;   // device word used by the streaming renderer to hand out jobs
};

int pick_step_tile(int64_t n, int forced);     // envs per CTA of a step launch over n envs (forced = 0: automatic)
cudaError_t launch_step(const BatchView &B, int ni, int forced_tile, const int32_t *actions, float *obs, float *reward,
                        uint8_t *done, uint8_t *ran, uint16_t *avail, cudaStream_t s);
cudaError_t launch_reset(const BatchView &B, int ni, const uint8_t *mask, float *obs, cudaStream_t s);
cudaError_t launch_primitive(const BatchView &B, int ni, const int32_t *actions, float *obs, float *reward, uint8_t *done, cudaStream_t s);
cudaError_t launch_init_with_state(const BatchView &B, int ni, const double *states, const uint8_t *mask, cudaStream_t s);
// observation rows of the envs [r_begin, r_begin + r_count) from their stored state
cudaError_t launch_obs(const BatchView &B, int ni, float *obs, cudaStream_t s);
// tg_step_frames: the M selected envs take one gym step tick by tick; per-tick states go to the snapshot batch S (M * T slots)
cudaError_t launch_trace(const BatchView &B, const BatchView &S, int ni, const int64_t *env_ids, int M, const int32_t *actions, int T,
                         int32_t *n_ticks, float *obs, float *rew, uint8_t *done, uint8_t *ran, cudaStream_t s);
cudaError_t launch_mask(const BatchView &B, int ni, uint8_t *mask, cudaStream_t s);
cudaError_t launch_get_state(const BatchView &B, const tg_state_view &v, cudaStream_t s);
cudaError_t launch_set_state(const BatchView &B, const tg_state_view &v, cudaStream_t s);
cudaError_t launch_render(const BatchView &B, const RenderView &R, int64_t first, int64_t count,
                          uint8_t *frames, cudaStream_t s, const int32_t *const *lists = nullptr, const int64_t *counts = nullptr);
cudaError_t launch_blend(const BatchView &B, const RenderView &R, int64_t first, int64_t n_surfaces, int64_t per,
                         uint8_t *surfaces, int alpha_objs, int alpha_player, cudaStream_t s);
cudaError_t launch_blit_alpha(uint8_t *target, int tw, int th, const uint8_t *source, int sw, int sh, int channels,
                              int x0, int y0, int opacity, cudaStream_t s);

}  // namespace tg
