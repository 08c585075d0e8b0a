// Kernel launchers shared between the translation units of libtreasure_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "tg_types.h"

namespace tg {

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

cudaError_t launch_step(const BatchView &B, int ni, const int32_t *actions, float *obs, float *reward,
                        uint8_t *done, uint8_t *ran, uint16_t *avail, cudaStream_t s);
cudaError_t launch_reset(const BatchView &B, int ni, const uint8_t *mask, float *obs, cudaStream_t s);
cudaError_t launch_primitive(const BatchView &B, int ni, const int32_t *actions, float *obs, float *reward, uint8_t *done, cudaStream_t s);
cudaError_t launch_init_with_state(const BatchView &B, int ni, const double *states, const uint8_t *mask, cudaStream_t s);
cudaError_t launch_mask(const BatchView &B, int ni, uint8_t *mask, cudaStream_t s);
cudaError_t launch_get_state(const BatchView &B, const tg_state_view &v, cudaStream_t s);
cudaError_t launch_set_state(const BatchView &B, const tg_state_view &v, cudaStream_t s);
cudaError_t launch_render(const BatchView &B, const RenderView &R, int64_t first, int64_t count,
                          uint8_t *frames, cudaStream_t s);
cudaError_t launch_blend(const BatchView &B, const RenderView &R, int64_t first, int64_t n_surfaces, int64_t per,
                         uint8_t *surfaces, int alpha_objs, int alpha_player, cudaStream_t s);
cudaError_t launch_blit_alpha(uint8_t *target, int tw, int th, const uint8_t *source, int sw, int sh, int channels,
                              int x0, int y0, int opacity, cudaStream_t s);
cudaError_t render_configure();   // one-time function attributes (dynamic shared memory opt-in)

}  // namespace tg
