// Analysis helpers of the reference drawer on the device:
//   _TreasureGameDrawer.blend       (_treasure_game_drawer.py:207-231)  -> tg_blend_kernel
//   _TreasureGameDrawer.blit_alpha  (:198-205)                          -> tg_blit_alpha_kernel
// (draw_background_to_surface :165-182 is the precomposed tile layer and draw_to_surface :184-196 has the
// pixels of draw_domain; both are served by the renderer in tg_render.cu.)
//
// blend() overlays the objects of one state (all but the handle bases, which go on at full opacity) with
// opacity alpha_objs and the hero with opacity alpha_player onto an existing surface; the research code
// calls it once per sampled state to visualise a set of states, so the batched form blends `per` consecutive
// envs, in order, onto each surface.  One CTA owns a band of 16 surface rows: the band is TMA-bulk-loaded into
// shared memory, every env of the surface is blended into it there (objects on a transparent RGBA overlay
// in shared memory, then overlay -> band, then the hero), and the band is TMA-bulk-stored once.
// Pixel arithmetic = pygame 1.9.6 / SDL 1.2 as restated in oracle/render_oracle.py (UNPINNED).
#include <cuda_runtime.h>
#include "tg_device.cuh"
#include "tg_launch.h"

namespace tg {

constexpr int BLEND_THREADS = 256;
constexpr int BLEND_ROWS = 16;

__device__ __forceinline__ uint32_t smem_a(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Canvas {
    uint8_t *rgb;       // band of the target surface, rows x W x 3
    uint32_t *ov;       // transparent overlay (new_surf, :209), rows x W, packed R | G<<8 | B<<16 | A<<24
    int W, H, y0, rows;
};

// SDL 1.2 per-pixel-alpha blit onto an opaque surface, one channel
__device__ __forceinline__ int px_alpha(int s, int d, int a) { return a == 0 ? d : a == 255 ? s : d + (((s - d) * a) >> 8); }
// SDL 1.2 per-surface alpha (set_alpha(a) on an opaque surface), one channel
__device__ __forceinline__ int surf_alpha(int s, int d, int a) {
    return a == 255 ? s : a == 128 ? (((s & 0xFE) + (d & 0xFE)) >> 1) + (s & d & 1) : d + (((s - d) * a) >> 8);
}

// draw.line / draw.circle on the SRCALPHA overlay: the mapped colour, alpha 255, no blending
__device__ __forceinline__ void ov_plot(const Canvas &c, int x, int y, uint32_t rgba) {
    if (x < 0 || x >= c.W || y < c.y0 || y >= c.y0 + c.rows || y >= c.H) return;
    c.ov[(y - c.y0) * c.W + x] = rgba;
}
// pygame 1.9.x draw.c drawline()
__device__ void ov_thin_line(const Canvas &c, int x1, int y1, int x2, int y2, uint32_t rgba) {
    int dx = x2 - x1, dy = y2 - y1;
    const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
    dx = sx * dx + 1; dy = sy * dy + 1;
    int x = x1, y = y1, err = 0;
    if (dx >= dy) {
        for (int i = 0; i < dx; i++) { ov_plot(c, x, y, rgba); x += sx; err += dy; if (err >= dx) { err -= dx; y += sy; } }
    } else {
        for (int i = 0; i < dy; i++) { ov_plot(c, x, y, rgba); y += sy; err += dx; if (err >= dy) { err -= dy; x += sx; } }
    }
}
__device__ void ov_span(const Canvas &c, int xa, int y, int xb, uint32_t rgba) {
    if (xa > xb) { const int t = xa; xa = xb; xb = t; }
    for (int x = xa; x <= xb; x++) ov_plot(c, x, y, rgba);
}
// pygame 1.9.x draw.circle(width=0) -> draw_fillellipse(x, y, r, r)
__device__ void ov_disc(const Canvas &c, int x, int y, int rad, uint32_t rgba) {
    if (rad <= 0) { ov_plot(c, x, y, rgba); return; }
    int oj = 0xFFFF, ok = 0xFFFF, ix = 0, iy = rad * 64, h, i, j, k;
    do {
        h = (ix + 8) >> 6; i = (iy + 8) >> 6;
        j = (h * rad) / rad; k = (i * rad) / rad;
        if (ok != k && oj != k && k < rad) { ov_span(c, x - h, y - k - 1, x + h - 1, rgba); ov_span(c, x - h, y + k, x + h - 1, rgba); ok = k; }
        if (oj != j && ok != j && k != j) { ov_span(c, x - i, y + j, x + i - 1, rgba); ov_span(c, x - i, y - j - 1, x + i - 1, rgba); oj = j; }
        ix = ix + iy / rad;
        iy = iy - ix / rad;
    } while (i > h);
}

// sprite rows that fall into the band: [r0, r1)
__device__ __forceinline__ bool band_rows(const Canvas &c, int oy, int &r0, int &r1) {
    r0 = max(c.y0 - oy, 0); r1 = min(min(c.y0 + c.rows, c.H) - oy, S);
    return r0 < r1;
}

// draw_object(obj, new_surf): per-pixel-alpha sprite onto the SRCALPHA overlay = pygame alphablit_alpha (ALPHA_BLEND)
__device__ void ov_blit48(const Canvas &c, const uint32_t *__restrict__ spr, int ox, int oy) {
    int r0, r1;
    if (!band_rows(c, oy, r0, r1)) return;
    for (int p = r0 * S + threadIdx.x; p < r1 * S; p += BLEND_THREADS) {
        const int sx = p % S, sy = p / S, x = ox + sx, y = oy + sy;
        if (x < 0 || x >= c.W) continue;
        const uint32_t s = __ldg(spr + p);
        uint32_t *dp = c.ov + (y - c.y0) * c.W + x;
        const uint32_t d = *dp;
        const int dA = d >> 24;
        if (dA == 0) { *dp = s; continue; }
        const int sA = s >> 24;
        uint32_t out = 0;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const int sc = (s >> (8 * ch)) & 255, dc = (d >> (8 * ch)) & 255;
            out |= (uint32_t)((((dc << 8) + (sc - dc) * sA + sc) >> 8) & 255) << (8 * ch);
        }
        out |= (uint32_t)(sA + dA - (sA * dA) / 255) << 24;
        *dp = out;
    }
}

// surf.blit(sprite) straight onto the target band (handle base, :220): SDL per-pixel alpha
__device__ void rgb_blit48(const Canvas &c, const uint32_t *__restrict__ spr, int ox, int oy) {
    int r0, r1;
    if (!band_rows(c, oy, r0, r1)) return;
    for (int p = r0 * S + threadIdx.x; p < r1 * S; p += BLEND_THREADS) {
        const int sx = p % S, sy = p / S, x = ox + sx, y = oy + sy;
        if (x < 0 || x >= c.W) continue;
        const uint32_t s = __ldg(spr + p);
        const int a = s >> 24;
        uint8_t *d = c.rgb + ((y - c.y0) * c.W + x) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) d[ch] = (uint8_t)px_alpha((s >> (8 * ch)) & 255, d[ch], a);
    }
}

// blit_alpha(surf, sprite, location, opacity) for a 48x48 sprite (:198-205): per pixel
//   temp = per-pixel-alpha blend of the sprite over the target, target = per-surface-alpha blend of temp over target
__device__ void rgb_blit48_alpha(const Canvas &c, const uint32_t *__restrict__ spr, int ox, int oy, int opacity) {
    int r0, r1;
    if (!band_rows(c, oy, r0, r1)) return;
    for (int p = r0 * S + threadIdx.x; p < r1 * S; p += BLEND_THREADS) {
        const int sx = p % S, sy = p / S, x = ox + sx, y = oy + sy;
        if (x < 0 || x >= c.W) continue;
        const uint32_t s = __ldg(spr + p);
        const int a = s >> 24;
        uint8_t *d = c.rgb + ((y - c.y0) * c.W + x) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const int dc = d[ch];
            d[ch] = (uint8_t)surf_alpha(px_alpha((s >> (8 * ch)) & 255, dc, a), dc, opacity);
        }
    }
}

__global__ void __launch_bounds__(BLEND_THREADS)
tg_blend_kernel(BatchView B, RenderView R, int64_t first, int64_t per, uint8_t *__restrict__ surfaces,
                int alpha_objs, int alpha_player) {
    extern __shared__ __align__(128) uint8_t blend_smem[];
    __shared__ uint64_t bar;
    const int W = R.frame_w, H = R.frame_h;
    Canvas c;
    c.W = W; c.H = H; c.y0 = blockIdx.x * BLEND_ROWS; c.rows = min(BLEND_ROWS, H - c.y0);
    c.rgb = blend_smem;
    c.ov = reinterpret_cast<uint32_t *>(blend_smem + (((size_t)BLEND_ROWS * W * 3 + 127) & ~(size_t)127));
    const uint32_t bytes = (uint32_t)(c.rows * W * 3);             // multiple of 16: W is a multiple of 48
    uint8_t *gband = surfaces + ((size_t)blockIdx.y * H + c.y0) * (size_t)W * 3;
    const uint32_t bar_a = smem_a(&bar), band_a = smem_a(c.rgb);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(band_a), "l"(gband), "r"(bytes), "r"(bar_a) : "memory");
    }
    __syncthreads();
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar_a) : "memory");
    } while (!ok);

    const int npx = c.rows * W;
    for (int64_t k = 0; k < per; k++) {
        const int64_t env = first + (int64_t)blockIdx.y * per + k;
        const int lid = B.level_id ? B.level_id[env] : 0;
        const LevelBlob &L = B.levels[lid];
        const uint32_t *spr = R.assets[lid].sprites;
        const uint4 cs = B.core[env];
        const uint32_t f = cs.y;
        uint32_t items[4] = {cs.z, cs.w, 0u, 0u};
        if (B.items23) { const uint2 h = B.items23[env]; items[2] = h.x; items[3] = h.y; }
        for (int p = threadIdx.x; p < npx; p += BLEND_THREADS) c.ov[p] = 0u;        // new_surf: transparent black (:209)
        __syncthreads();
        for (int o = 0; o < L.n_objs; o++) {                                         // :211, file order
            const int kind = L.obj_kind[o], i = L.obj_idx[o];
            if (kind == TG_HANDLE) {                                                 // :212-220
                const int ox = L.handle_cx[i] * S, oy = L.handle_cy[i] * S;
                if (threadIdx.x == 0) {                                              // plots clip to the band
                    const double ang = B.angles[(int64_t)i * B.n + env];
                    const double th = __dadd_rn(__dmul_rn(1.5707963267948966, ang), 0.7853981633974483);
                    const int ex = (int)__dadd_rn((double)(ox + S / 2), __dmul_rn(36.0, cos(th)));
                    const int ey = (int)__dsub_rn((double)(oy + S), __dmul_rn(36.0, sin(th)));
                    const uint32_t line = 47u | (79u << 8) | (79u << 16) | (255u << 24), disc = 255u | (255u << 24);
                    const bool shallow = abs(ox + S / 2 - ex) > abs(oy + S - ey);      // clip_and_draw_line_width, width 5
                    const int xi = shallow ? 0 : 1, yi = shallow ? 1 : 0;
                    for (int w = -2; w <= 2; w++) ov_thin_line(c, ox + S / 2 + xi * w, oy + S + yi * w, ex + xi * w, ey + yi * w, line);
                    ov_disc(c, ex, ey, S / 10, disc);
                }
                rgb_blit48(c, spr + TG_SPR_HANDLE_BASE * S * S, ox, oy);
            } else if (kind == TG_DOOR) {
                const bool closed = (f >> (F_DOORS + i)) & 1u;
                ov_blit48(c, spr + (closed ? TG_SPR_DOOR_CLOSED : TG_SPR_DOOR_OPEN) * S * S, L.door_cx[i] * S, L.door_cy[i] * S);
            } else if (kind == TG_BOLT) {
                const bool locked = (f >> (F_BOLTS + i)) & 1u;
                ov_blit48(c, spr + (locked ? TG_SPR_BOLT_LOCKED : TG_SPR_BOLT_OPEN) * S * S, L.bolt_cx[i] * S, L.bolt_cy[i] * S);
            } else {                                                                 // key / gold; skipped when x < 0 (:240-241)
                const int ox = lo16(items[i]), oy = hi16(items[i]);
                if (ox < 0) continue;
                ov_blit48(c, spr + (kind == TG_KEY ? TG_SPR_KEY : TG_SPR_GOLD) * S * S, ox, oy);
            }
            __syncthreads();
        }
        // blit_alpha(surf, new_surf, (0, 0), int(255 * alpha_objs))  (:223)
        for (int p = threadIdx.x; p < npx; p += BLEND_THREADS) {
            const uint32_t s = c.ov[p];
            const int a = s >> 24;
            if (a == 0) continue;                                   // temp == target there: every blend returns target
            uint8_t *d = c.rgb + p * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                const int dc = d[ch];
                d[ch] = (uint8_t)surf_alpha(px_alpha((s >> (8 * ch)) & 255, dc, a), dc, alpha_objs);
            }
        }
        __syncthreads();
        // hero, flipped when facing left (:225-231)
        rgb_blit48_alpha(c, spr + ((f & 1u) ? TG_SPR_HERO_RIGHT : TG_SPR_HERO_LEFT) * S * S, core_px(cs.x) - S / 2, hi16(cs.x), alpha_player);
        __syncthreads();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gband), "r"(band_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// blit_alpha(target, source, location, opacity) (:198-205) for any RGB / RGBA source: one thread per source pixel.
__global__ void tg_blit_alpha_kernel(uint8_t *__restrict__ target, int tw, int th, const uint8_t *__restrict__ source,
                                     int sw, int sh, int channels, int x0, int y0, int opacity) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= sw * sh) return;
    const int sx = p % sw, sy = p / sw, x = x0 + sx, y = y0 + sy;
    if (x < 0 || x >= tw || y < 0 || y >= th) return;              // the final blit is clipped to the target
    const uint8_t *s = source + (size_t)p * channels;
    const int a = channels == 4 ? s[3] : 255;
    uint8_t *d = target + ((size_t)y * tw + x) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const int dc = d[ch];
        d[ch] = (uint8_t)surf_alpha(px_alpha(s[ch], dc, a), dc, opacity);
    }
}

static size_t blend_smem_bytes(int W) { return (((size_t)BLEND_ROWS * W * 3 + 127) & ~(size_t)127) + (size_t)BLEND_ROWS * W * 4; }

cudaError_t launch_blend(const BatchView &B, const RenderView &R, int64_t first, int64_t n_surfaces, int64_t per,
                         uint8_t *surfaces, int alpha_objs, int alpha_player, cudaStream_t s) {
    static size_t allowed[MAX_DEVICES] = {};                       // per device: function attributes are per device
    const size_t smem = blend_smem_bytes(R.frame_w);
    const int dslot = device_slot();
    if (smem > 48 * 1024 && smem > allowed[dslot]) {
        cudaError_t e = cudaFuncSetAttribute(tg_blend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        allowed[dslot] = smem;
    }
    const unsigned bands = (unsigned)((R.frame_h + BLEND_ROWS - 1) / BLEND_ROWS);
    for (int64_t s0 = 0; s0 < n_surfaces; s0 += 32768) {           // gridDim.y limit
        const unsigned ny = (unsigned)((n_surfaces - s0 < 32768) ? n_surfaces - s0 : 32768);
        tg_blend_kernel<<<dim3(bands, ny), BLEND_THREADS, smem, s>>>(B, R, first + s0 * per, per,
                                                                      surfaces + (size_t)s0 * R.frame_h * R.frame_w * 3,
                                                                      alpha_objs, alpha_player);
    }
    return cudaGetLastError();
}

cudaError_t launch_blit_alpha(uint8_t *target, int tw, int th, const uint8_t *source, int sw, int sh, int channels,
                              int x0, int y0, int opacity, cudaStream_t s) {
    const int n = sw * sh;
    tg_blit_alpha_kernel<<<(n + 255) / 256, 256, 0, s>>>(target, tw, th, source, sw, sh, channels, x0, y0, opacity);
    return cudaGetLastError();
}

}  // namespace tg
