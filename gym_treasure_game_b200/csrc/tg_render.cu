// RGB observation renderer: replaces TreasureGame.render('rgb_array')
// (treasure_game.py:98-104) -> _TreasureGameDrawer.draw_domain (_treasure_game_drawer.py:136-163)
// -> draw_object (:238-269).
//
// The tile layer of draw_domain is a constant image per level (the drawer reseeds its private
// Random with 12 on every frame, :137), precomposed on the host.  One CTA renders one 48-row
// band (one cell row) of one frame:
//   TMA bulk load  background band (L2-resident) -> shared memory        (UBLKCP, mbarrier)
//   patch the dynamic layer in shared memory: doors, key, gold, bolt, lever + handle base, hero
//   TMA bulk store shared memory -> the env's frame in HBM                (UBLKCP, bulk_group)
// so every frame byte is written to HBM exactly once with full-line bulk writes and nothing is
// read from HBM but 48 bytes of env state (background and sprites are served from L2).
#include <cuda_runtime.h>
#include "tg_device.cuh"
#include "tg_launch.h"

namespace tg {

constexpr int RENDER_THREADS = 256;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Band {
    uint8_t *px;      // shared memory, 48 rows x W x 3
    int W, H, y0;     // frame size, first frame row of this band
};

// Surface.blit of a 48x48 per-pixel-alpha sprite onto the opaque screen.  Blend as SDL 1.2's
// 32-bit ARGB->RGB blitters (MMX and generic forms): alpha 0 keeps dst, 255 copies src, otherwise
// d + (((s - d) * a) >> 8) per channel (floor).  See DESIGN.md "Render semantics".
__device__ __forceinline__ void blit48(const Band &b, const uint32_t *__restrict__ spr, int ox, int oy) {
    if (oy + S <= b.y0 || oy >= b.y0 + S) return;
    const int r0 = max(b.y0 - oy, 0), r1 = min(b.y0 + S - oy, S);      // sprite rows inside the band
    for (int p = r0 * S + threadIdx.x; p < r1 * S; p += RENDER_THREADS) {
        const int sx = p % S, sy = p / S;
        const int x = ox + sx, y = oy + sy;
        if (x < 0 || x >= b.W || y >= b.H) continue;
        const uint32_t s = __ldg(spr + p);
        const int a = s >> 24;
        if (a == 0) continue;
        uint8_t *d = b.px + ((y - b.y0) * b.W + x) * 3;
        const int sr = s & 255, sg = (s >> 8) & 255, sb = (s >> 16) & 255;
        if (a == 255) { d[0] = (uint8_t)sr; d[1] = (uint8_t)sg; d[2] = (uint8_t)sb; }
        else {
            const int dr = d[0], dg = d[1], db = d[2];
            d[0] = (uint8_t)(dr + (((sr - dr) * a) >> 8));
            d[1] = (uint8_t)(dg + (((sg - dg) * a) >> 8));
            d[2] = (uint8_t)(db + (((sb - db) * a) >> 8));
        }
    }
}

__device__ __forceinline__ void plot(const Band &b, int x, int y, uint8_t r, uint8_t g, uint8_t bl) {
    if (x < 0 || x >= b.W || y < b.y0 || y >= b.y0 + S || y >= b.H) return;
    uint8_t *d = b.px + ((y - b.y0) * b.W + x) * 3;
    d[0] = r; d[1] = g; d[2] = bl;
}

// pygame 1.9.x draw.c drawline(): error-accumulating line over max(|dx|,|dy|)+1 pixels
__device__ void thin_line(const Band &b, int x1, int y1, int x2, int y2, uint8_t r, uint8_t g, uint8_t bl) {
    int dx = x2 - x1, dy = y2 - y1;
    const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
    dx = sx * dx + 1; dy = sy * dy + 1;
    int x = x1, y = y1, err = 0;
    if (dx >= dy) {
        for (int i = 0; i < dx; i++) { plot(b, x, y, r, g, bl); x += sx; err += dy; if (err >= dx) { err -= dx; y += sy; } }
    } else {
        for (int i = 0; i < dy; i++) { plot(b, x, y, r, g, bl); y += sy; err += dx; if (err >= dy) { err -= dy; x += sx; } }
    }
}

// pygame 1.9.x draw.line(width=5): the thin line plus copies shifted by +-1, +-2 along x for
// steep lines and along y for shallow ones (clip_and_draw_line_width)
__device__ void thick_line5(const Band &b, int x1, int y1, int x2, int y2, uint8_t r, uint8_t g, uint8_t bl) {
    const bool shallow = abs(x1 - x2) > abs(y1 - y2);
    const int xi = shallow ? 0 : 1, yi = shallow ? 1 : 0;
    for (int k = -2; k <= 2; k++) thin_line(b, x1 + xi * k, y1 + yi * k, x2 + xi * k, y2 + yi * k, r, g, bl);
}

__device__ void hspan(const Band &b, int xa, int y, int xb, uint8_t r, uint8_t g, uint8_t bl) {
    if (xa > xb) { int t = xa; xa = xb; xb = t; }
    for (int x = xa; x <= xb; x++) plot(b, x, y, r, g, bl);
}

// pygame 1.9.x draw.circle(width=0) -> draw_fillellipse(x, y, r, r)
__device__ void fill_circle(const Band &b, int x, int y, int rad, uint8_t r, uint8_t g, uint8_t bl) {
    if (rad <= 0) { plot(b, x, y, r, g, bl); return; }
    int oh = 0xFFFF, oi = 0xFFFF, oj = 0xFFFF, ok = 0xFFFF;
    (void)oh; (void)oi;
    int ix = 0, iy = rad * 64, h, i, j, k;
    do {
        h = (ix + 8) >> 6; i = (iy + 8) >> 6;
        j = (h * rad) / rad; k = (i * rad) / rad;
        if (ok != k && oj != k && k < rad) {
            hspan(b, x - h, y - k - 1, x + h - 1, r, g, bl);
            hspan(b, x - h, y + k, x + h - 1, r, g, bl);
            ok = k;
        }
        if (oj != j && ok != j && k != j) {
            hspan(b, x - i, y + j, x + i - 1, r, g, bl);
            hspan(b, x - i, y - j - 1, x + i - 1, r, g, bl);
            oj = j;
        }
        ix = ix + iy / rad;
        iy = iy - ix / rad;
    } while (i > h);
}

__global__ void __launch_bounds__(RENDER_THREADS)
tg_render_kernel(BatchView B, RenderView R, int64_t first, uint8_t *__restrict__ frames) {
    extern __shared__ __align__(128) uint8_t band_px[];
    __shared__ uint64_t bar;
    const int bandi = blockIdx.x;
    const int64_t env = first + blockIdx.y;
    const int lid = B.level_id ? B.level_id[env] : 0;
    const LevelBlob &L = B.levels[lid];
    const RenderAssets A = R.assets[lid];
    const int W = R.frame_w, H = R.frame_h;
    const uint32_t band_bytes = (uint32_t)(S * W * 3);
    const uint32_t bar_a = smem_addr(&bar), band_a = smem_addr(band_px);

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(band_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(band_a), "l"(A.background + (size_t)bandi * band_bytes), "r"(band_bytes), "r"(bar_a) : "memory");
    }
    // env state (uniform across the CTA) while the band is in flight
    const uint4 c = B.core[env];
    const uint32_t f = c.y;
    const int px = lo16(c.x), py = hi16(c.x);
    uint32_t items[4] = {c.z, c.w, 0u, 0u};
    if (B.items23) { const uint2 h = B.items23[env]; items[2] = h.x; items[3] = h.y; }
    __syncthreads();
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar_a) : "memory");
    } while (!ok);

    Band b; b.px = band_px; b.W = W; b.H = H; b.y0 = bandi * S;
    const uint32_t *spr = A.sprites;
    // objects in file order (drawer.py:154-155)
    for (int o = 0; o < L.n_objs; o++) {
        const int kind = L.obj_kind[o], i = L.obj_idx[o];
        if (kind == TG_DOOR) {                                                       // :243-247
            const int ox = L.door_cx[i] * S, oy = L.door_cy[i] * S;
            if (oy + S <= b.y0 || oy >= b.y0 + S) continue;
            const bool closed = (f >> (F_DOORS + i)) & 1u;
            blit48(b, spr + (closed ? TG_SPR_DOOR_CLOSED : TG_SPR_DOOR_OPEN) * S * S, ox, oy);
            __syncthreads();
        } else if (kind == TG_KEY || kind == TG_GOLD) {                              // :248-251
            const int ox = lo16(items[i]), oy = hi16(items[i]);
            if (ox < 0) continue;                                                    // :240-241
            if (oy + S <= b.y0 || oy >= b.y0 + S) continue;
            blit48(b, spr + (kind == TG_KEY ? TG_SPR_KEY : TG_SPR_GOLD) * S * S, ox, oy);
            __syncthreads();
        } else if (kind == TG_BOLT) {                                                // :252-256
            const int ox = L.bolt_cx[i] * S, oy = L.bolt_cy[i] * S;
            if (oy + S <= b.y0 || oy >= b.y0 + S) continue;
            const bool locked = (f >> (F_BOLTS + i)) & 1u;
            blit48(b, spr + (locked ? TG_SPR_BOLT_LOCKED : TG_SPR_BOLT_OPEN) * S * S, ox, oy);
            __syncthreads();
        } else if (kind == TG_HANDLE) {                                              // :257-266
            const int ox = L.handle_cx[i] * S, oy = L.handle_cy[i] * S;
            // lever reaches from row oy+48+2 up to about oy+48-36-4-2; base sprite covers the cell
            if (oy + S + 3 <= b.y0 || oy - 8 >= b.y0 + S) continue;
            if (threadIdx.x == 0) {
                const double ang = B.angles[(int64_t)i * B.n + env];
                const double th = __dadd_rn(__dmul_rn(1.5707963267948966, ang), 0.7853981633974483);
                const double r = 36.0;                                               // yscale * 0.75
                const double sx = (double)(ox + S / 2), sy = (double)(oy + S);
                const int ex = (int)__dadd_rn(sx, __dmul_rn(r, cos(th)));
                const int ey = (int)__dsub_rn(sy, __dmul_rn(r, sin(th)));
                thick_line5(b, ox + S / 2, oy + S, ex, ey, 47, 79, 79);
                fill_circle(b, ex, ey, S / 10, 255, 0, 0);
            }
            __syncthreads();
            blit48(b, spr + TG_SPR_HANDLE_BASE * S * S, ox, oy);
            __syncthreads();
        }
    }
    // hero (drawer.py:157-161): flipped sprite when facing left, at (playerx - 24, playery)
    blit48(b, spr + ((f & 1u) ? TG_SPR_HERO_RIGHT : TG_SPR_HERO_LEFT) * S * S, px - S / 2, py);

    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
        uint8_t *dst = frames + ((size_t)blockIdx.y * H + (size_t)b.y0) * (size_t)W * 3;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(band_a), "r"(band_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

static bool g_render_configured = false;

cudaError_t render_configure() {
    if (g_render_configured) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(tg_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) g_render_configured = true;
    return e;
}

cudaError_t launch_render(const BatchView &B, const RenderView &R, int64_t first, int64_t count,
                          uint8_t *frames, cudaStream_t s) {
    cudaError_t e = render_configure();
    if (e != cudaSuccess) return e;
    const size_t band_bytes = (size_t)S * R.frame_w * 3;
    // grid.y is limited to 65535: render in slabs
    for (int64_t off = 0; off < count; off += 32768) {
        const int64_t c = (count - off < 32768) ? count - off : 32768;
        dim3 grid((unsigned)R.ch, (unsigned)c);
        tg_render_kernel<<<grid, RENDER_THREADS, band_bytes, s>>>(
            B, R, first + off, frames + (size_t)off * R.frame_h * R.frame_w * 3);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace tg
