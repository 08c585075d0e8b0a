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
#include <cstdlib>
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
tg_render_band_kernel(BatchView B, RenderView R, int64_t first, uint8_t *__restrict__ frames) {
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
    const int px = core_px(c.x), py = hi16(c.x);
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


// ===========================================================================
// v2: streaming renderer (single-level batches).  Persistent CTAs; a job is (unit type t, block of
// RS_EB environments) where a unit is UR consecutive frame rows.  For the whole job the pristine
// background rows of type t stay in shared memory (P, one TMA bulk load per job); units without
// dynamic content are TMA-stored to the frame straight from P, units with sprites are composed
// in one of two working copies (W0/W1: P -> W copy, per-pixel patches, TMA store).  Nothing is read
// from L2/HBM per unit but 16-40 bytes of env state, every frame byte is written exactly once by a
// bulk store, and several stores are in flight per SM.
// Patches need no barriers: a thread owns the screen columns x = tid (mod RS_THREADS), so all
// objects touching a pixel are applied by the same thread in file order; the lever (pygame
// draw.line width 5 + filled circle) is evaluated per pixel in closed form.
// ===========================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_PATCH_THREADS = RS_THREADS - 32;     // warps 1..7 compose dirty units; warp 0 streams clean ones
constexpr int RS_EB_MAX = 256;        // envs per job: array bound; the launcher picks eb <= RS_EB_MAX

__device__ __forceinline__ void patch_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(RS_PATCH_THREADS) : "memory"); }

__device__ __forceinline__ void wait_bulk_read(int pending) {
    switch (pending) {
    case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
    default: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
    }
}

struct Unit {
    uint8_t *buf;     // UR rows x W x 3 in shared memory
    int W, y0, UR;
};

__device__ __forceinline__ void put_px(const Unit &u, int x, int y, int r, int g, int b) {
    uint8_t *d = u.buf + ((y - u.y0) * u.W + x) * 3;
    d[0] = (uint8_t)r; d[1] = (uint8_t)g; d[2] = (uint8_t)b;
}

// Pixel ownership among the 224 composer threads: owner(x, y) = (x + 48 * (y & 3)) mod 224.  Every
// object touching a pixel is therefore applied by the same thread, in file order, with no barriers
// between objects, and a 48-wide sprite spreads over 192 threads (2 rows each in an 8-row unit)
// instead of 48 threads with 8 serial rows each.
__device__ __forceinline__ void blit_cols(const Unit &u, const uint32_t *__restrict__ spr, int ox, int oy) {
    const int r0 = max(u.y0 - oy, 0), r1 = min(u.y0 + u.UR - oy, S);
    if (r0 >= r1) return;
    int v = ((int)threadIdx.x - 32 - ox) % RS_PATCH_THREADS;       // patch threads are tid 32..255
    if (v < 0) v += RS_PATCH_THREADS;
    if (v >= 4 * S) return;
    const int cls = v / S, sx = v - cls * S, x = ox + sx;
    if (x < 0 || x >= u.W) return;
    int sy = r0 + ((cls - (oy + r0)) & 3);                         // first row >= r0 with ((oy + sy) & 3) == cls
    for (; sy < r1; sy += 16) {                                    // up to 4 rows of this class in flight
        uint32_t px4[4];
#pragma unroll
        for (int q = 0; q < 4; q++) px4[q] = (sy + 4 * q < r1) ? __ldg(spr + (sy + 4 * q) * S + sx) : 0u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t s = px4[q];
            const int a = s >> 24;
            if (a == 0) continue;                                  // also skips rows beyond r1
            uint8_t *d = u.buf + ((oy + sy + 4 * q - u.y0) * u.W + x) * 3;
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
}

// Lever of one handle (drawer.py:258-265): pygame draw.line width 5 in (47,79,79), then a filled circle
// of radius 4 in (255,0,0) at the end point.  One thread per plotted pixel: thread q < 5*dmax plots pixel
// i = q % dmax of copy k = q / dmax - 2 of drawline() (x-major: (ax + sx*i, ay + sy*floor(i*dyp/dxp)),
// y-major: roles swapped; clip_and_draw_line_width() shifts the copies along x when |dx| <= |dy|, else
// along y); the last 64 threads plot the disc spans.  Line pixels under the disc are skipped, so the two
// groups never write the same pixel and no barrier is needed between them.  The caller brackets this
// with __syncthreads() because pixel ownership differs from the column-owned sprite blits.
__device__ __forceinline__ void lever_pixels(const Unit &u, int x1, int y1, int ex, int ey, int rad,
                                             const int8_t *disc_lo, const int8_t *disc_hi) {
    auto inside = [&](int x, int y) { return x >= 0 && x < u.W && y >= u.y0 && y < u.y0 + u.UR; };
    auto in_disc = [&](int x, int y) {
        const int row = y - ey + rad;
        return row >= 0 && row < 2 * rad && x - ex >= disc_lo[row] && x - ex <= disc_hi[row];
    };
    int dx = ex - x1, dy = ey - y1;
    const int sx = dx < 0 ? -1 : 1, sy = dy < 0 ? -1 : 1;
    const bool shift_y = abs(dx) > abs(dy);
    const int dxp = sx * dx + 1, dyp = sy * dy + 1;
    const int dmax = max(dxp, dyp);
    const int nline = 5 * dmax, ndisc = 2 * rad * 16;              // disc: 2*rad rows x up to 16 px
    for (int q = (int)threadIdx.x - 32; q < nline + ndisc; q += RS_PATCH_THREADS) {
        if (q < nline) {
            const int k = q / dmax - 2, i = q % dmax;
            const int ax = x1 + (shift_y ? 0 : k), ay = y1 + (shift_y ? k : 0);
            int x, y;
            if (dxp >= dyp) { x = ax + sx * i; y = ay + sy * ((i * dyp) / dxp); }
            else { y = ay + sy * i; x = ax + sx * ((i * dxp) / dyp); }
            if (inside(x, y) && !in_disc(x, y)) put_px(u, x, y, 47, 79, 79);
        } else {
            const int r = (q - nline) >> 4, c = (q - nline) & 15;
            const int x = ex + disc_lo[r] + c, y = ey - rad + r;
            if (x <= ex + disc_hi[r] && inside(x, y)) put_px(u, x, y, 255, 0, 0);
        }
    }
}

// column ownership over n threads (my = this thread's index among them)
__device__ __forceinline__ int owned_col_n(int ox, int my, int n) {
    int sx = (my - ox) % n;
    if (sx < 0) sx += n;
    return sx < S ? sx : -1;
}

// blit_cols with an explicit owner set (used by all 256 threads when the static prefix is baked)
__device__ __forceinline__ void blit_cols_n(const Unit &u, const uint32_t *__restrict__ spr, int ox, int oy, int my, int n) {
    const int r0 = max(u.y0 - oy, 0), r1 = min(u.y0 + u.UR - oy, S);
    if (r0 >= r1) return;
    const int sx = owned_col_n(ox, my, n);
    if (sx < 0) return;
    const int x = ox + sx;
    if (x < 0 || x >= u.W) return;
    for (int rb = r0; rb < r1; rb += 8) {
        uint32_t px8[8];
#pragma unroll
        for (int q = 0; q < 8; q++) px8[q] = (rb + q < r1) ? __ldg(spr + (rb + q) * S + sx) : 0u;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint32_t s = px8[q];
            const int a = s >> 24;
            if (a == 0) continue;
            uint8_t *d = u.buf + ((oy + rb + q - u.y0) * u.W + x) * 3;
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
}

__device__ __forceinline__ void mbar_wait(uint32_t bar_a, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar_a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load(uint32_t dst_a, const void *src, uint32_t bytes, uint32_t bar_a) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_a), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
}
__device__ __forceinline__ void tma_store(void *dst, uint32_t src_a, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_a), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Per job the unit type (rows y0 .. y0+UR) is fixed, so the objects that can touch it are known.  The
// leading run (file order) of two-state objects among them -- doors, bolts, keys/gold resting at their
// initial cell -- is the *static prefix*: it is blitted once per job into C (the job's copy of the
// pristine rows) in the state combination most envs of the job share (key K*).  Then
//   mode 0: env's prefix state == K* and nothing else touches the rows  -> streamed straight from C
//   mode 1: == K* but a lever / handle base / moved item / the hero touches the rows -> W = copy of C + those
//   mode 2: prefix state differs from K*  -> W = pristine rows re-fetched by TMA (L2) + every object
// Drawing order is preserved: everything outside the prefix comes later in file order, the hero last.
__global__ void __launch_bounds__(RS_THREADS)
tg_render_stream_kernel(BatchView B, RenderAssets A, int lid, const int32_t *__restrict__ env_list, int W, int H, int UR, int EB,
                        int64_t first, int64_t count, uint8_t *__restrict__ frames, unsigned long long *__restrict__ job_counter, int dbg) {
    extern __shared__ __align__(128) uint8_t rs_smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint4 s_core[RS_EB_MAX];
    __shared__ uint2 s_items23[RS_EB_MAX];
    __shared__ short2 s_lever[RS_EB_MAX][TG_MAX_HANDLES];
    __shared__ int8_t disc_lo[32], disc_hi[32];
    __shared__ short4 s_obj[TG_MAX_OBJECTS];          // draw list: x = ox, y = oy, z = kind, w = index in kind
    __shared__ short2 s_item_init[TG_MAX_ITEMS];      // initial pixel position of every key / gold
    __shared__ int s_nobj, s_nhandles, s_nitems, s_votes;
    __shared__ uint16_t s_key[RS_EB_MAX];
    __shared__ uint8_t s_mode[RS_EB_MAX];
    __shared__ int32_t s_off[RS_EB_MAX];              // frame index of the job's envs (env - first): consecutive, or from env_list
    __shared__ long long s_job;
    __shared__ uint16_t s_torder[320];               // unit types in job order: the ones a lever touches first
    const int tid = threadIdx.x;
    const uint32_t unit_bytes = (uint32_t)(UR * W * 3);
    uint8_t *C = rs_smem, *W0 = rs_smem + unit_bytes, *W1 = rs_smem + 2 * (size_t)unit_bytes;
    const uint32_t bar_a = smem_addr(&bar), bar2_a = smem_addr(&bar2);
    const LevelBlob &L = B.levels[lid];               // one layout per launch: a mixed batch is rendered level by level (env_list)
    const uint32_t *spr = A.sprites;
    const int rad = S / 10;                                       // int(xscale / 10), drawer.py:265

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_nobj = L.n_objs; s_nhandles = L.n_handles; s_nitems = L.n_items;
        for (int o = 0; o < L.n_objs; o++) {
            const int kind = L.obj_kind[o], i = L.obj_idx[o];
            int cx = 0, cy = 0;
            if (kind == TG_DOOR) { cx = L.door_cx[i]; cy = L.door_cy[i]; }
            else if (kind == TG_BOLT) { cx = L.bolt_cx[i]; cy = L.bolt_cy[i]; }
            else if (kind == TG_HANDLE) { cx = L.handle_cx[i]; cy = L.handle_cy[i]; }
            else { cx = L.item_cx[i]; cy = L.item_cy[i]; s_item_init[i] = make_short2((short)(cx * S), (short)(cy * S)); }
            s_obj[o] = make_short4((short)(cx * S), (short)(cy * S), (short)kind, (short)i);
        }
        {   // unit types in job order: lever rows spread evenly among the others, none of them among the last types
            const int nt = H / UR;
            int nl = 0;
            auto is_lever = [&](int t) {
                bool lever = false;
                for (int hnd = 0; hnd < L.n_handles; hnd++) {
                    const int oy = L.handle_cy[hnd] * S;
                    lever = lever || (oy + S + 3 > t * UR && oy < t * UR + UR);
                }
                return lever;
            };
            for (int t = 0; t < nt; t++) nl += is_lever(t) ? 1 : 0;
            const int every = nl ? max(1, (nt - nl - 4) / nl) : nt;         // clean types between two lever types
            int k = 0, tl = 0, tc = 0, run = every;                           // next lever / clean type to place
            while (k < nt) {
                while (tl < nt && !is_lever(tl)) tl++;
                while (tc < nt && is_lever(tc)) tc++;
                if (tl < nt && (run >= every || tc >= nt)) { s_torder[k++] = (uint16_t)tl++; run = 0; }
                else { s_torder[k++] = (uint16_t)tc++; run++; }
            }
        }
        // filled-circle spans of pygame 1.9.x draw_fillellipse(rx = ry = rad), recorded per row
        for (int i = 0; i < 2 * rad; i++) { disc_lo[i] = 127; disc_hi[i] = -128; }
        int oj = 0xFFFF, ok = 0xFFFF, ix = 0, iy = rad * 64, h, i, j, k;
        auto span = [&](int xa, int yrel, int xb) {
            if (xa > xb) { int t = xa; xa = xb; xb = t; }
            const int row = yrel + rad;
            if (row < 0 || row >= 2 * rad) return;
            disc_lo[row] = (int8_t)min((int)disc_lo[row], xa); disc_hi[row] = (int8_t)max((int)disc_hi[row], xb);
        };
        do {
            h = (ix + 8) >> 6; i = (iy + 8) >> 6; j = (h * rad) / rad; k = (i * rad) / rad;
            if (ok != k && oj != k && k < rad) { span(-h, -k - 1, h - 1); span(-h, k, h - 1); ok = k; }
            if (oj != j && ok != j && k != j) { span(-i, j, i - 1); span(-i, -j - 1, i - 1); oj = j; }
            ix = ix + iy / rad; iy = iy - ix / rad;
        } while (i > h);
    }
    __syncthreads();

    const int ntypes = H / UR;
    const int nblocks = (int)((count + EB - 1) / EB);
    const int64_t njobs = (int64_t)ntypes * nblocks;
    const int nobj = s_nobj;
    uint32_t parity = 0, parity2 = 0;
    int G = 0, lastW[2] = {-1, -1}, nd = 0;          // bulk-group bookkeeping of the composer's issuer (tid 32)

    for (;;) {
        // jobs are handed out dynamically (lever rows cost more than plain rows); the previous job's final
        // __syncthreads() guarantees nobody still reads s_job
        if (tid == 0) s_job = (long long)atomicAdd(job_counter, 1ull);
        __syncthreads();
        const int64_t job = s_job;
        if (job >= njobs) break;
        // unit type major, types permuted: a lever-row job composes every unit (two working copies = 64 KB of stores in flight
        // per CTA against the streamer's seven), so lever rows are spread among the plain rows instead of having every
        // resident CTA in them at the same time.
        const int t = s_torder[(int)(job / nblocks)], blk = (int)(job % nblocks);
        const int64_t e0 = (int64_t)blk * EB;
        const int ne = (int)min((int64_t)EB, count - e0);
        const int y0 = t * UR;
        const uint8_t *pristine = A.background + (size_t)y0 * W * 3;
        auto rows_hit = [&](int oy, int above, int below) { return oy + S + below > y0 && oy - above < y0 + UR; };
        if (tid == 0) {
            wait_bulk_read(0);                       // the streamer's stores out of C have drained
            tma_load(smem_addr(C), pristine, unit_bytes, bar_a);
            s_votes = 0;
        }
        if (tid < ne) {
            const int64_t env = env_list ? (int64_t)env_list[e0 + tid] : first + e0 + tid;
            s_off[tid] = (int32_t)(env - first);
            s_core[tid] = B.core[env];
            s_items23[tid] = B.items23 ? B.items23[env] : make_uint2(0u, 0u);
            for (int hnd = 0; hnd < s_nhandles; hnd++) {                          // drawer.py:258-262
                const double ang = B.angles[(int64_t)hnd * B.n + env];
                const double th = __dadd_rn(__dmul_rn(1.5707963267948966, ang), 0.7853981633974483);
                const double sxd = (double)(L.handle_cx[hnd] * S + S / 2), syd = (double)(L.handle_cy[hnd] * S + S);
                s_lever[tid][hnd] = make_short2((short)(int)__dadd_rn(sxd, __dmul_rn(36.0, cos(th))),
                                                (short)(int)__dsub_rn(syd, __dmul_rn(36.0, sin(th))));
            }
        }
        // static prefix of this unit type (uniform): bit o of pre_mask = object o is baked into C
        uint32_t pre_mask = 0;
        for (int o = 0; o < nobj; o++) {
            const short4 ob = s_obj[o];
            if (ob.z == TG_HANDLE) { if (rows_hit(ob.y, 0, 3)) break; else continue; }
            if (rows_hit(ob.y, 0, 0)) pre_mask |= 1u << o;
        }
        __syncthreads();                             // state block visible; s_votes reset
        // prefix state of every env: door closed / bolt locked / item resting at its initial cell
        bool item_moved_here = false;                // an item away from its initial cell touches these rows
        if (tid < ne) {
            const uint4 c = s_core[tid];
            const uint32_t items[4] = {c.z, c.w, s_items23[tid].x, s_items23[tid].y};
            uint32_t key = 0;
            for (int o = 0; o < nobj; o++) {
                const short4 ob = s_obj[o];
                bool bit = false;
                if (ob.z == TG_DOOR) bit = (c.y >> (F_DOORS + ob.w)) & 1u;
                else if (ob.z == TG_BOLT) bit = (c.y >> (F_BOLTS + ob.w)) & 1u;
                else if (ob.z != TG_HANDLE) {
                    uint32_t it = items[0];
#pragma unroll
                    for (int q = 1; q < TG_MAX_ITEMS; q++) if (q == ob.w) it = items[q];
                    bit = lo16(it) == ob.x && hi16(it) == ob.y;
                    if (!bit && lo16(it) >= 0 && rows_hit(hi16(it), 0, 0)) item_moved_here = true;
                }
                if (bit && ((pre_mask >> o) & 1u)) key |= 1u << o;
            }
            s_key[tid] = (uint16_t)key;
        }
        __syncthreads();
        if (tid < ne) {                              // majority vote for K*
            int cnt = 0;
            const uint16_t mine = s_key[tid];
            for (int v = 0; v < ne; v++) cnt += (s_key[v] == mine);
            atomicMax(&s_votes, (cnt << 16) | tid);
        }
        __syncthreads();
        const uint32_t kstar = s_key[s_votes & 0xFFFF];
        mbar_wait(bar_a, parity);                    // pristine rows have landed in C
        parity ^= 1u;
        {   // bake the prefix in state K* into C (all 256 threads, column-owned, file order)
            Unit u; u.buf = C; u.W = W; u.y0 = y0; u.UR = UR;
            for (int o = 0; o < nobj; o++) {
                if (!((pre_mask >> o) & 1u)) continue;
                const short4 ob = s_obj[o];
                const bool bit = (kstar >> o) & 1u;
                const uint32_t *sp = nullptr;
                if (ob.z == TG_DOOR) sp = spr + (bit ? TG_SPR_DOOR_CLOSED : TG_SPR_DOOR_OPEN) * S * S;
                else if (ob.z == TG_BOLT) sp = spr + (bit ? TG_SPR_BOLT_LOCKED : TG_SPR_BOLT_OPEN) * S * S;
                else if (bit) sp = spr + (ob.z == TG_KEY ? TG_SPR_KEY : TG_SPR_GOLD) * S * S;
                if (sp) blit_cols_n(u, sp, ob.x, ob.y, tid, RS_THREADS);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tid < ne) {
            const uint4 c = s_core[tid];
            bool dyn = rows_hit(hi16(c.x), 0, 0) || item_moved_here;           // hero, moved items
            for (int o = 0; o < nobj; o++) {                                     // objects outside the prefix
                const short4 ob = s_obj[o];
                if ((pre_mask >> o) & 1u) continue;
                if (ob.z == TG_HANDLE) dyn |= rows_hit(ob.y, 0, 3);
                else if (ob.z == TG_DOOR || ob.z == TG_BOLT) dyn |= rows_hit(ob.y, 0, 0);
                // items outside the prefix: resting ones count if their cell touches the rows, moved ones are in item_moved_here
                else {
                    const uint32_t items[4] = {c.z, c.w, s_items23[tid].x, s_items23[tid].y};
                    uint32_t it = items[0];
#pragma unroll
                    for (int q = 1; q < TG_MAX_ITEMS; q++) if (q == ob.w) it = items[q];
                    dyn |= lo16(it) >= 0 && rows_hit(hi16(it), 0, 0);
                }
            }
            const bool same = s_key[tid] == (uint16_t)kstar && !item_moved_here;
            s_mode[tid] = (dbg == 1) ? 0 : (!same ? 2 : (dyn ? 1 : 0));
        }
        __syncthreads();
        uint8_t *row_dst = frames + (size_t)y0 * (size_t)W * 3;       // + frame index * frame_bytes
        const size_t frame_bytes = (size_t)H * W * 3;

        if (tid < 32) {
            // ---- role 1 (warp 0): stream the units that are exactly C ----
            if (tid == 0) {
                for (int uidx = 0; uidx < ne; uidx++) {
                    if (s_mode[uidx] != 0) continue;
                    wait_bulk_read(6);
                    tma_store(row_dst + (size_t)s_off[uidx] * frame_bytes, smem_addr(C), unit_bytes);
                }
            }
        } else {
            // ---- role 2 (warps 1..7): compose the other units in W0/W1 and store them ----
            const int pt = tid - 32;
            for (int uidx = 0; uidx < ne; uidx++) {
                const int mode = s_mode[uidx];
                if (mode == 0) continue;
                const uint4 c = s_core[uidx];
                const uint32_t f = c.y;
                const int px = core_px(c.x), py = hi16(c.x);
                const uint32_t items[4] = {c.z, c.w, s_items23[uidx].x, s_items23[uidx].y};
                const int kbuf = nd & 1;
                nd++;
                uint8_t *wbuf = kbuf ? W1 : W0;
                if (pt == 0) {
                    if (lastW[kbuf] >= 0) wait_bulk_read(min(G - 1 - lastW[kbuf], 6));
                    if (mode == 2) tma_load(smem_addr(wbuf), pristine, unit_bytes, bar2_a);
                }
                patch_barrier();
                if (mode == 2) {
                    mbar_wait(bar2_a, parity2);
                } else if (dbg != 2) {               // working copy <- C
                    const uint4 *src4 = reinterpret_cast<const uint4 *>(C);
                    uint4 *dst4 = reinterpret_cast<uint4 *>(wbuf);
                    for (uint32_t q = pt; q < unit_bytes / 16; q += RS_PATCH_THREADS) dst4[q] = src4[q];
                }
                if (mode == 2) parity2 ^= 1u;
                patch_barrier();
                Unit u; u.buf = wbuf; u.W = W; u.y0 = y0; u.UR = UR;
                for (int o = 0; o < nobj; o++) {                                       // drawer.py:154-155
                    if (mode == 1 && ((pre_mask >> o) & 1u)) continue;                 // already in C
                    const short4 ob = s_obj[o];
                    const int kind = ob.z, i = ob.w;
                    if (kind == TG_DOOR) {
                        const bool closed = (f >> (F_DOORS + i)) & 1u;
                        blit_cols(u, spr + (closed ? TG_SPR_DOOR_CLOSED : TG_SPR_DOOR_OPEN) * S * S, ob.x, ob.y);
                    } else if (kind == TG_KEY || kind == TG_GOLD) {
                        uint32_t it = items[0];
#pragma unroll
                        for (int q = 1; q < TG_MAX_ITEMS; q++) if (q == i) it = items[q];
                        if (lo16(it) >= 0)                                             // drawer.py:240-241
                            blit_cols(u, spr + (kind == TG_KEY ? TG_SPR_KEY : TG_SPR_GOLD) * S * S, lo16(it), hi16(it));
                    } else if (kind == TG_BOLT) {
                        const bool locked = (f >> (F_BOLTS + i)) & 1u;
                        blit_cols(u, spr + (locked ? TG_SPR_BOLT_LOCKED : TG_SPR_BOLT_OPEN) * S * S, ob.x, ob.y);
                    } else {
                        if (rows_hit(ob.y, 0, 3)) {    // lever rows: oy+8 .. oy+50 (uniform branch)
                            patch_barrier();
                            if (dbg != 3) lever_pixels(u, ob.x + S / 2, ob.y + S, s_lever[uidx][i].x, s_lever[uidx][i].y, rad, disc_lo, disc_hi);
                            patch_barrier();
                        }
                        blit_cols(u, spr + TG_SPR_HANDLE_BASE * S * S, ob.x, ob.y);
                    }
                }
                blit_cols(u, spr + ((f & 1u) ? TG_SPR_HERO_RIGHT : TG_SPR_HERO_LEFT) * S * S, px - S / 2, py);   // drawer.py:157-161
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                patch_barrier();
                if (pt == 0) {
                    wait_bulk_read(6);
                    tma_store(row_dst + (size_t)s_off[uidx] * frame_bytes, smem_addr(wbuf), unit_bytes);
                    lastW[kbuf] = G;
                    G++;
                }
            }
        }
        __syncthreads();                             // both roles are done with C before the next job reloads it
    }
    if (tid == 32) wait_bulk_read(0);                // shared memory must outlive the last bulk stores' reads
    if (tid == 0) wait_bulk_read(0);
}

// dynamic shared memory opt-in, once per device (function attributes are per device)
static cudaError_t render_configure() {
    static bool configured[MAX_DEVICES] = {};
    const int d = device_slot();
    if (configured[d]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(tg_render_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tg_render_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e == cudaSuccess) configured[d] = true;
    return e;
}

// rows per unit of the streaming renderer: the largest divisor of 48 whose three buffers fit twice per SM.
// Measured on B200 (final kernel, dynamic jobs, 16384 frames of 672x624, 256 envs per job):
// UR 8/12/16/24 -> 5.07/5.04/5.61/4.46 TB/s.
static int pick_unit_rows(int W) {
    static int forced = -1;
    if (forced < 0) { const char *v = getenv("TG_RENDER_UR"); forced = v ? atoi(v) : 0; }
    if (forced > 0 && 48 % forced == 0) return forced;
    static const int cand[] = {48, 24, 16, 12, 8, 6, 4};
    for (int ur : cand) if ((size_t)3 * ur * W * 3 <= 100 * 1024) return ur;
    return 4;
}

// one launch of the streaming renderer: `count` envs of layout `lid`, consecutive from `first` or listed in env_list
static cudaError_t launch_stream(const BatchView &B, const RenderView &R, int lid, const int32_t *env_list, int64_t first, int64_t count,
                                 uint8_t *frames, cudaStream_t s) {
    const int ur = pick_unit_rows(R.frame_w);
    const size_t smem = (size_t)3 * ur * R.frame_w * 3;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tg_render_stream_kernel, RS_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    // envs per job: as many as possible (the per-job set-up -- pristine rows, vote, baking -- is amortised
    // over them: 16384 frames, 32/64/128/256 envs -> 5.27/5.38/5.49/5.61 TB/s) while leaving >= 8 jobs per
    // resident CTA so that the dynamic job queue still balances (4096 frames: 256 -> 4.3, 64 -> 5.45 TB/s)
    static int eb_forced = -1;
    if (eb_forced < 0) { const char *v = getenv("TG_RENDER_EB"); eb_forced = v ? atoi(v) : 0; }
    const int64_t slots = (int64_t)device_sm_count() * per_sm;
    int eb = RS_EB_MAX;
    while (eb > 16 && (int64_t)(R.frame_h / ur) * ((count + eb - 1) / eb) < 8 * slots) eb >>= 1;
    if (eb_forced >= 16 && eb_forced <= RS_EB_MAX) eb = eb_forced;
    const int64_t njobs = (int64_t)(R.frame_h / ur) * ((count + eb - 1) / eb);
    const int64_t grid = njobs < slots ? njobs : slots;
    static int dbg = -1;
    if (dbg < 0) { const char *v = getenv("TG_RENDER_DBG"); dbg = v ? atoi(v) : 0; }
    e = cudaMemsetAsync(R.job_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    tg_render_stream_kernel<<<(unsigned)grid, RS_THREADS, smem, s>>>(B, R.assets[lid], lid, env_list, R.frame_w, R.frame_h, ur, eb, first, count,
                                                                     frames, R.job_counter, dbg);
    return cudaGetLastError();
}

// Frames of the envs [first, first + count).  Single-layout batches: one launch of the streaming renderer.  Mixed
// batches: one launch per layout over that layout's envs of the range (lists[l] = their indices, DEV, ascending;
// counts[l] of them) -- each launch keeps one layout's pristine rows, prefix and sprites, exactly like a single-layout
// batch (the per-band kernel it replaces re-read the background per frame: 2.95 TB/s).  Without lists (the snapshot
// batches of tg_step_frames) mixed batches use the per-band kernel.
cudaError_t launch_render(const BatchView &B, const RenderView &R, int64_t first, int64_t count,
                          uint8_t *frames, cudaStream_t s, const int32_t *const *lists, const int64_t *counts) {
    cudaError_t e = render_configure();
    if (e != cudaSuccess) return e;
    static int force_v1 = -1;
    if (force_v1 < 0) { const char *v = getenv("TG_RENDER_V1"); force_v1 = (v && v[0] == '1') ? 1 : 0; }
    if (B.n_levels == 1 && !force_v1) return launch_stream(B, R, 0, nullptr, first, count, frames, s);
    if (lists && counts && !force_v1) {
        for (int l = 0; l < B.n_levels; l++) {
            if (counts[l] <= 0) continue;
            e = launch_stream(B, R, l, lists[l], first, counts[l], frames, s);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    const size_t band_bytes = (size_t)S * R.frame_w * 3;
    // grid.y is limited to 65535: render in slabs
    for (int64_t off = 0; off < count; off += 32768) {
        const int64_t c = (count - off < 32768) ? count - off : 32768;
        dim3 grid((unsigned)R.ch, (unsigned)c);
        tg_render_band_kernel<<<grid, RENDER_THREADS, band_bytes, s>>>(
            B, R, first + off, frames + (size_t)off * R.frame_h * R.frame_w * 3);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace tg
