// K1 -- fused front-end, generation 4: BGR -> grey -> 3x3 blur -> the two adaptive-threshold decisions -> bit masks in ONE
// kernel; no intermediate plane ever reaches HBM.  Replaces /root/reference/ysmr/track_eval.py:180 (cv2.cvtColor), :182
// (cv2.GaussianBlur) and :189-208 (two cv2.adaptiveThreshold calls) for the adaptive modes.
//
// Exact bound-and-refine.  cv2.adaptiveThreshold decides  b - rint(mean) > t  with `mean` the 11x11 float32 Gaussian of the
// blurred image.  That Gaussian costs ~22 dependent-order FP32 operations per pixel and made the previous generation
// issue-bound at 0.18 of the HBM roofline.  But almost every pixel is background whose decision is not close: it only
// needs a LOWER BOUND L <= mean with  b - t - 1/2 < L.  The bound used here is itself a separable filter, in exact integer
// arithmetic on byte dot products (IDP.4A, 4 multiply-adds per instruction):
//
//     mean(x, y) = sum_ij k_i k_j b(x+i, y+j)  >=  base + sum_ij h_i h_j (b(x+i, y+j) - base)          (*)
//
// for any  0 <= h <= k  pointwise and any  base <= min b  over the window.  h is chosen so that ONE filter output serves a
// 2 x 2 group of pixels: h_j = floor(256 min(k_j, k_j+1)) / 256 on the 10 taps the windows of x and x+1 share (mass 0.80 per
// axis), so the bound costs 1.5 + 0.75 IDP per pixel instead of 22 FMAs.  What (*) gives away is (1 - 0.64) of the mean of
// (b - base), i.e. ~1.5 grey levels on a noisy background with base = the tile minimum -- against a decision distance of 5.
// Pixels the bound cannot clear ("candidates": all foreground, its rim and a few noise peaks, 0.15 % of a cfg2 frame) are
// collected in a shared-memory list and decided by the exact OpenCV-order float32 arithmetic (front_arith.cuh, the same
// functions the generation-3 kernels use), including OpenCV's scalar-tail columns.  Masks are therefore bit-identical to
// cv2's whatever the bound does; a loose bound only costs time.  (scripts/bound_proto.py is the numpy model of the bound;
// tests/test_gpu_detect.py holds the mask parity tests.)
//
// Dark-on-light (THRESH_BINARY_INV) runs the same bound on complemented bytes p = 255 - b (negated weights and a different
// start value of the dot-product chain, no extra instruction).
//
// One CTA = one tile of tw x th pixels of one frame:
//   1a  BGR (or grey) words -> Q15 luma (two IDP.2A per pixel) -> grey tile in shared memory, REFLECT_101 at the image border
//   1b  3x3 binomial blur: horizontal pass as byte dot products, vertical pass sliding through registers -> blurred tile;
//       tile minimum / maximum on the way
//   1c  BORDER_REPLICATE margins of the blurred tile (tiles at the image border only)
//   2a  row pass of the bound (3 IDP per pixel pair), results as bytes, transposed
//   2b  column pass (3 IDP per 2 x 2 group) -> per-group threshold byte -> candidate test -> candidate list
//   3   exact float32 Gaussian + both decisions for the candidates -> bits into the shared mask tiles
//   4   mask tiles -> HBM (the only global stores of the kernel: 2 bits per pixel)
#include "frontend.cuh"
#include "front_arith.cuh"

namespace ysmr {

constexpr int FT_THREADS = 256;

struct FusedGeom {
    int tw, th;                 // tile size: tw % 32 == 0, th % 4 == 0, tw * th <= 65536
    int tiles_x, tiles_y;
    int gp;                     // byte pitch of the grey / blurred tiles: >= tw + 24, gp % 8 == 0, (gp / 8) odd
    int rp;                     // byte pitch of the transposed row-pass bytes: >= th + 8, rp % 4 == 0, (rp / 4) odd
    int t_q;                    // polarised decision threshold (see fused_geometry)
    int off_blur, off_mask, off_list, off_misc;   // shared-memory offsets (bytes); grey tile and row-pass bytes at 0
    int smem_bytes;
};

// weights of the bound: h_j = floor(256 * min(k_j, k_j+1)), taps at offsets -4 .. +5 relative to the even pixel of a pair
// k = cv2.getGaussianKernel(11, 0, CV_32F): 256 k = 2.2559, 6.9488, 16.669, 31.142, 45.312, 51.345
constexpr int BH0 = 2, BH1 = 6, BH2 = 16, BH3 = 31, BH4 = 45;
constexpr int BH_SUM = 2 * (BH0 + BH1 + BH2 + BH3 + BH4);                // 200
__host__ __device__ constexpr uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}
// taps 0..9 = (H0 H1 H2 H3 H4 H4 H3 H2 H1 H0).  A pair whose first tap sits at byte 0 of a word uses (KA, KB, KC); a pair
// whose first tap sits at byte 2 uses (KA2, KB2, KC2).
constexpr uint32_t KA = pack4(BH0, BH1, BH2, BH3), KB = pack4(BH4, BH4, BH3, BH2), KC = pack4(BH1, BH0, 0, 0);
constexpr uint32_t KA2 = pack4(0, 0, BH0, BH1), KB2 = pack4(BH2, BH3, BH4, BH4), KC2 = pack4(BH3, BH2, BH1, BH0);
constexpr uint32_t KAn = pack4(-BH0, -BH1, -BH2, -BH3), KBn = pack4(-BH4, -BH4, -BH3, -BH2), KCn = pack4(-BH1, -BH0, 0, 0);
constexpr uint32_t KA2n = pack4(0, 0, -BH0, -BH1), KB2n = pack4(-BH2, -BH3, -BH4, -BH4), KC2n = pack4(-BH3, -BH2, -BH1, -BH0);

struct FusedMisc {              // small shared block
    uint32_t tmin, tmax;        // minimum / maximum of the blurred tile
    uint32_t count;             // candidates
};

__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ uint32_t lds16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// 4 x (unsigned byte of a) * (signed byte of b) + c  (IDP.4A.U8.S8)
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t vmin3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t vmax3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }

// horizontal 1-2-1 of eight grey pixels (words w0, w1) with their neighbours (byte 3 of wl, byte 0 of wr): four packed
// pairs (h0,h1) .. (h6,h7), each value + 2 so that the vertical 1-2-1 carries the rounding constant 8 of (sum + 8) >> 4
struct HRow4 {
    uint32_t h01, h23, h45, h67;
};
__device__ __forceinline__ HRow4 hblur8(uint32_t wl, uint32_t w0, uint32_t w1, uint32_t wr)
{
    const uint32_t a0 = __byte_perm(wl, w0, 0x6543);              // (g-1, g0, g1, g2)
    const uint32_t b0 = __byte_perm(w0, w1, 0x4321);              // (g1, g2, g3, g4)
    const uint32_t a1 = __byte_perm(w0, w1, 0x6543);              // (g3, g4, g5, g6)
    const uint32_t b1 = __byte_perm(w1, wr, 0x4321);              // (g5, g6, g7, g8)
    constexpr uint32_t LO = 0x00010201u, HI = 0x01020100u;        // weights (1,2,1,0) and (0,1,2,1)
    HRow4 h;
    h.h01 = __dp4a(a0, LO, 2u) + (__dp4a(a0, HI, 2u) << 16);
    h.h23 = __dp4a(b0, LO, 2u) + (__dp4a(b0, HI, 2u) << 16);
    h.h45 = __dp4a(a1, LO, 2u) + (__dp4a(a1, HI, 2u) << 16);
    h.h67 = __dp4a(b1, LO, 2u) + (__dp4a(b1, HI, 2u) << 16);
    return h;
}
__device__ __forceinline__ HRow4 hadd(const HRow4 &a, const HRow4 &b)
{
    HRow4 s;
    s.h01 = a.h01 + b.h01; s.h23 = a.h23 + b.h23; s.h45 = a.h45 + b.h45; s.h67 = a.h67 + b.h67;
    return s;
}

template <int C>
__device__ __forceinline__ uint32_t grey_word_slow(const uint8_t *frame, int w, int gy, int vx)
{
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) v |= grey_px<C>(frame, w, gy, reflect101(vx + k, w)) << (8 * k);
    return v;
}

extern __shared__ __align__(16) unsigned char ysmr_fused_smem[];

template <int C>
__global__ void __launch_bounds__(FT_THREADS, 3) fused_front_kernel(FrontParams p, FusedGeom g)
{
    const int tid = threadIdx.x, lane = tid & 31;
    // tile of this CTA
    int bid = blockIdx.x;
    const int tix = bid % g.tiles_x; bid /= g.tiles_x;
    const int tiy = bid % g.tiles_y;
    const int f = bid / g.tiles_y;
    const int tw = g.tw, th = g.th, gp = g.gp, rp = g.rp;
    const int tx0 = tix * tw, ty0 = tiy * th;
    const int W = p.w, H = p.h;
    const uint8_t *frame = p.frames + (int64_t)f * p.frame_stride;

    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(ysmr_fused_smem);
    const uint32_t s_grey = s_base, s_rq = s_base, s_blur = s_base + g.off_blur, s_list = s_base + g.off_list;
    uint32_t *smask = reinterpret_cast<uint32_t *>(ysmr_fused_smem + g.off_mask);
    FusedMisc *misc = reinterpret_cast<FusedMisc *>(ysmr_fused_smem + g.off_misc);
    const int mw = tw >> 5;                                       // mask words per tile row
    const int n_mask = th * mw;

    for (int i = tid; i < 2 * n_mask; i += FT_THREADS) smask[i] = 0u;
    if (tid == 0) { misc->tmin = 0xFFFFu; misc->tmax = 0u; misc->count = 0u; }

    // ---- 1a: grey tile.  Row r <-> virtual image row ty0 - 6 + r, byte c <-> virtual column tx0 - 12 + c; virtual positions
    // outside the image hold the REFLECT_101 pixel (what cv2.GaussianBlur reads there).
    {
        const int gwr = (tw + 24) >> 2;                           // words per row
        const int n_rows = th + 12;
        const int dr = FT_THREADS / gwr, dc = FT_THREADS - dr * gwr;
        int r = tid / gwr, c4 = tid - r * gwr;
        constexpr int U = 4;
        while (r < n_rows) {
            uint32_t raw[U][C == 3 ? 3 : 1];
            int rr[U], cc[U];
            bool fast[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                rr[u] = r; cc[u] = c4;
                const int vx = tx0 - 12 + 4 * c4;
                fast[u] = r < n_rows && vx >= 0 && vx + 3 < W;
                if (fast[u]) {
                    const int gy = reflect101(ty0 - 6 + r, H);
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(frame + ((int64_t)gy * W + vx) * C);
                    raw[u][0] = __ldg(q);
                    if (C == 3) { raw[u][1] = __ldg(q + 1); raw[u][2] = __ldg(q + 2); }
                }
                c4 += dc; r += dr;
                if (c4 >= gwr) { c4 -= gwr; ++r; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (rr[u] >= n_rows) break;
                uint32_t v;
                if (fast[u]) v = C == 3 ? grey4_of_bgr(raw[u][0], raw[u][1], raw[u][2]) : raw[u][0];
                else v = grey_word_slow<C>(frame, W, reflect101(ty0 - 6 + rr[u], H), tx0 - 12 + 4 * cc[u]);
                sts32(s_grey + rr[u] * gp + 4 * cc[u], v);
            }
        }
    }
    __syncthreads();

    // ---- 1b: blurred tile, same geometry: row rb <-> virtual row ty0 - 5 + rb = grey rows rb, rb+1, rb+2.  A thread owns 8
    // columns (blurred words 2cg+1, 2cg+2) of a chunk of rows; the vertical 1-2-1 slides through registers.
    {
        const int ncg = (tw + 16) >> 3;
        const int nchunk = FT_THREADS / ncg;
        const int n_rows = th + 10;
        const int rpc = (n_rows + nchunk - 1) / nchunk;
        const int chunk = tid / ncg, cg = tid - chunk * ncg;
        uint32_t mn0 = 0xFFFFFFFFu, mx0 = 0u;
        if (chunk < nchunk) {
            const int r0 = chunk * rpc, r1 = min(n_rows, r0 + rpc);
            if (r0 < r1) {
                uint32_t ga = s_grey + r0 * gp + 8 * cg;          // grey words 2cg .. 2cg+3 of grey row r0
                uint32_t ba = s_blur + r0 * gp + 8 * cg + 4;
                auto hrow = [&](uint32_t a) {
                    const uint2 lo = lds64(a), hi = lds64(a + 8);
                    return hblur8(lo.x, lo.y, hi.x, hi.y);
                };
                HRow4 b = hrow(ga + gp);                          // h(y)
                HRow4 s_prev = hadd(hrow(ga), b);                 // h(y-1) + h(y)
                ga += 2 * gp;
                for (int r = r0; r < r1; ++r) {
                    const HRow4 c = hrow(ga);                     // h(y+1)
                    const HRow4 s_cur = hadd(b, c);
                    const HRow4 v = hadd(s_prev, s_cur);          // 16 * blurred + rounding, per 16-bit lane
                    mn0 = vmin3_u16x2(mn0, v.h01, v.h23); mn0 = vmin3_u16x2(mn0, v.h45, v.h67);
                    mx0 = vmax3_u16x2(mx0, v.h01, v.h23); mx0 = vmax3_u16x2(mx0, v.h45, v.h67);
                    sts32(ba, __byte_perm(v.h01 >> 4, v.h23 >> 4, 0x6420));
                    sts32(ba + 4, __byte_perm(v.h45 >> 4, v.h67 >> 4, 0x6420));
                    s_prev = s_cur; b = c;
                    ga += gp; ba += gp;
                }
            }
        }
        uint32_t mn = min(mn0 & 0xFFFFu, mn0 >> 16) >> 4, mx = max(mx0 & 0xFFFFu, mx0 >> 16) >> 4;
        mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) { atomicMin(&misc->tmin, mn); atomicMax(&misc->tmax, mx); }
    }
    __syncthreads();

    // optional dumps of the production planes (parity tests): image pixels of this tile only
    if (p.dbg_grey || p.dbg_blurred) {
        const int64_t plane = (int64_t)H * W;
        for (int i = tid; i < tw * th; i += FT_THREADS) {
            const int yy = i / tw, xx = i - yy * tw;
            const int x = tx0 + xx, y = ty0 + yy;
            if (x < W && y < H) {
                if (p.dbg_grey) p.dbg_grey[f * plane + (int64_t)y * W + x] = ysmr_fused_smem[(yy + 6) * gp + 12 + xx];
                if (p.dbg_blurred) p.dbg_blurred[f * plane + (int64_t)y * W + x] = ysmr_fused_smem[g.off_blur + (yy + 5) * gp + 12 + xx];
            }
        }
        __syncthreads();
    }

    // ---- 1c: BORDER_REPLICATE margins of the blurred tile (what cv2.adaptiveThreshold's Gaussian sees outside the image)
    {
        const int n_rows = th + 10;
        const bool fix_l = tx0 - 8 < 0, fix_r = tx0 + tw + 8 > W;
        const bool fix_t = ty0 - 5 < 0, fix_b = ty0 + th + 5 > H;
        if (fix_l || fix_r) {
            // word granularity: W % 4 == 0 and tx0 % 4 == 0, so a word is entirely inside or outside
            const int c_first = 12 - tx0;                                         // byte column of image column 0 (left tiles)
            const int c_last = 12 + (W - 1 - tx0);                                // byte column of image column W-1
            const int wl_n = fix_l ? 2 : 0;                                       // words 1, 2 (columns -8 .. -1)
            const int w_r0 = (c_last + 1) >> 2;                                   // first word right of the image
            const int wr_n = fix_r ? max(0, ((tw + 20) >> 2) - w_r0) : 0;
            const int per_row = wl_n + wr_n;
            for (int i = tid; i < n_rows * per_row; i += FT_THREADS) {
                const int r = i / per_row, k = i - r * per_row;
                const uint32_t row = s_blur + r * gp;
                if (k < wl_n) {
                    const uint32_t v = lds32(row + c_first) & 0xFFu;
                    sts32(row + 4 + 4 * k, v * 0x01010101u);
                } else {
                    const uint32_t v = lds32(row + c_last - 3) >> 24;
                    sts32(row + 4 * (w_r0 + k - wl_n), v * 0x01010101u);
                }
            }
            __syncthreads();
        }
        if (fix_t || fix_b) {
            const int wpr = (tw + 24) >> 2;
            const int r_first = 5 - ty0;                                          // tile row of image row 0 (top tiles)
            const int r_last = 5 + (H - 1 - ty0);                                 // tile row of image row H-1
            const int nt = fix_t ? r_first : 0;                                   // rows above the image
            const int nb = fix_b ? max(0, n_rows - 1 - r_last) : 0;
            for (int i = tid; i < (nt + nb) * wpr; i += FT_THREADS) {
                const int k = i / wpr, wd = i - k * wpr;
                const int dst = k < nt ? k : r_last + 1 + (k - nt);
                const int src = k < nt ? r_first : r_last;
                sts32(s_blur + dst * gp + 4 * wd, lds32(s_blur + src * gp + 4 * wd));
            }
            __syncthreads();
        }
    }

    // ---- polarity and scale of the bound
    const bool inv = p.inverted != 0;
    const int tmin = (int)misc->tmin, tmax = (int)misc->tmax;
    const int range = tmax - tmin;
    int sh = 0;
    while (((BH_SUM * range) >> sh) > 255) ++sh;                  // row-pass results must fit a byte
    const int base_p = inv ? 255 - tmax : tmin;                   // minimum of the polarised bytes
    const int acc_row = inv ? BH_SUM * tmax : -BH_SUM * tmin;     // start of the row chain: sum h (p - base_p) >= 0
    const int ka = inv ? (int)KAn : (int)KA, kb = inv ? (int)KBn : (int)KB, kc = inv ? (int)KCn : (int)KC;
    const int ka2 = inv ? (int)KA2n : (int)KA2, kb2 = inv ? (int)KB2n : (int)KB2, kc2 = inv ? (int)KC2n : (int)KC2;

    // ---- 2a: row pass.  Task (q, ry): blurred tile row ry + 1 (image row ty0 - 4 + ry), pixels 8q .. 8q+7 -> four pair results
    // rq[X][ry], X = 4q .. 4q+3, stored transposed (bytes along ry).  Lanes run along ry: conflict-free LDS.64 (gp / 8 odd)
    // and byte stores into consecutive bytes.
    {
        const int nry = th + 8, nq = tw >> 3;
        const int n_tasks = nq * nry;
        const int dq = FT_THREADS / nry, dry = FT_THREADS - dq * nry;
        int q = tid / nry, ry = tid - q * nry;
        for (int t = tid; t < n_tasks; t += FT_THREADS) {
            const uint32_t a = s_blur + (ry + 1) * gp + 8 * q + 8;                // relative words 2q-1 .. 2q+2
            const uint2 lo = lds64(a), hi = lds64(a + 8);
            int r0 = dp4a_us(lo.x, ka, acc_row); r0 = dp4a_us(lo.y, kb, r0); r0 = dp4a_us(hi.x, kc, r0);
            int r1 = dp4a_us(lo.x, ka2, acc_row); r1 = dp4a_us(lo.y, kb2, r1); r1 = dp4a_us(hi.x, kc2, r1);
            int r2 = dp4a_us(lo.y, ka, acc_row); r2 = dp4a_us(hi.x, kb, r2); r2 = dp4a_us(hi.y, kc, r2);
            int r3 = dp4a_us(lo.y, ka2, acc_row); r3 = dp4a_us(hi.x, kb2, r3); r3 = dp4a_us(hi.y, kc2, r3);
            const uint32_t o = s_rq + (4 * q) * rp + ry;
            sts8(o, (uint32_t)(r0 >> sh)); sts8(o + rp, (uint32_t)(r1 >> sh));
            sts8(o + 2 * rp, (uint32_t)(r2 >> sh)); sts8(o + 3 * rp, (uint32_t)(r3 >> sh));
            ry += dry; q += dq;
            if (ry >= nry) { ry -= nry; ++q; }
        }
    }
    __syncthreads();

    // ---- 2b: column pass + candidate test.  Task (X, m): pixel pair X, output rows 4m .. 4m+3 (two row pairs).  A pixel is
    // certain background iff  p <= floor(base_p + t_q + 0.48 + L), L = (column sum << sh) / 65536.
    {
        const int nx = tw >> 1, nm = th >> 2;
        const int c0 = ((base_p + g.t_q) * 65536 + 31457) >> sh;  // arithmetic shift: floor, i.e. towards "candidate"
        const int ts = 16 - sh;
        const int n_tasks = nx * nm;
        const int dm = FT_THREADS / nx, dX = FT_THREADS - dm * nx;
        int m = tid / nx, X = tid - m * nx;
        for (int t = tid; t < n_tasks; t += FT_THREADS) {
            const int m_ = m, X_ = X;
            X += dX; m += dm;
            if (X >= nx) { X -= nx; ++m; }
            const uint32_t a = s_rq + X_ * rp + 4 * m_;
            const uint32_t w0 = lds32(a), w1 = lds32(a + 4), w2 = lds32(a + 8);
            int la = dp4a_us(w0, (int)KA, c0); la = dp4a_us(w1, (int)KB, la); la = dp4a_us(w2, (int)KC, la);
            int lb = dp4a_us(w0, (int)KA2, c0); lb = dp4a_us(w1, (int)KB2, lb); lb = dp4a_us(w2, (int)KC2, lb);
            const int ta = la >> ts, tb = lb >> ts;
            const int x = tx0 + 2 * X_;
            if (x >= W) continue;                                 // (W is even: a pair is inside or outside)
            const uint32_t ba = s_blur + (4 * m_ + 5) * gp + 12 + 2 * X_;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t v = lds16(ba + k * gp);
                int p0 = (int)(v & 0xFFu), p1 = (int)(v >> 8);
                if (inv) { p0 = 255 - p0; p1 = 255 - p1; }
                const int tt = k < 2 ? ta : tb;
                if ((p0 > tt || p1 > tt) && ty0 + 4 * m_ + k < H) {
                    if (p0 > tt) { const uint32_t idx = atomicAdd(&misc->count, 1u); asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_list + 2 * idx), "r"((4 * m_ + k) * tw + 2 * X_)); }
                    if (p1 > tt) { const uint32_t idx = atomicAdd(&misc->count, 1u); asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_list + 2 * idx), "r"((4 * m_ + k) * tw + 2 * X_ + 1)); }
                }
            }
        }
    }
    __syncthreads();

    // ---- 3: exact decisions of the candidates (OpenCV's float32 arithmetic, front_arith.cuh)
    {
        const int n_cand = (int)misc->count;
        for (int i = tid; i < n_cand; i += FT_THREADS) {
            uint32_t idx;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(idx) : "r"(s_list + 2 * i));
            const int yy = (int)idx / tw, xx = (int)idx - yy * tw;
            const int x = tx0 + xx;
            const bool row_tail = x >= p.row_tail_from, col_tail = x >= p.col_tail_from;
            const uint32_t a0 = s_blur + yy * gp + 12 + xx - 5;   // tile row yy + 5 - 5, byte column of x - 5
            const uint32_t al = a0 & ~3u;
            const uint32_t sel = 0x3210u + 0x1111u * (a0 & 3u);
            float r[11];
            int bc = 0;
#pragma unroll
            for (int j = 0; j < 11; ++j) {
                const uint32_t ra = al + j * gp;
                const uint32_t u0 = lds32(ra), u1 = lds32(ra + 4), u2 = lds32(ra + 8), u3 = lds32(ra + 12);
                const uint32_t v0 = __byte_perm(u0, u1, sel), v1 = __byte_perm(u1, u2, sel), v2 = __byte_perm(u2, u3, sel);
                float a[11];
                a[0] = (float)(v0 & 0xFFu); a[1] = (float)((v0 >> 8) & 0xFFu); a[2] = (float)((v0 >> 16) & 0xFFu); a[3] = (float)(v0 >> 24);
                a[4] = (float)(v1 & 0xFFu); a[5] = (float)((v1 >> 8) & 0xFFu); a[6] = (float)((v1 >> 16) & 0xFFu); a[7] = (float)(v1 >> 24);
                a[8] = (float)(v2 & 0xFFu); a[9] = (float)((v2 >> 8) & 0xFFu); a[10] = (float)((v2 >> 16) & 0xFFu);
                if (j == 5) bc = (int)((v1 >> 8) & 0xFFu);
                r[j] = gauss_row<true>(a, row_tail);
            }
            const float acc = gauss_col<true>(r[5], r[4], r[6], r[3], r[7], r[2], r[8], r[1], r[9], r[0], r[10], col_tail);
            int mean = __float2int_rn(acc);
            mean = mean < 0 ? 0 : (mean > 255 ? 255 : mean);
            const int d = bc - mean;
            const bool m_mask = (d > p.t_mask) != inv, m_mark = (d > p.t_marker) != inv;
            const int wi = yy * mw + (xx >> 5);
            const uint32_t bit = 1u << (xx & 31);
            if (m_mask) atomicOr(&smask[wi], bit);
            if (m_mark && p.marker_bits) atomicOr(&smask[n_mask + wi], bit);
        }
    }
    __syncthreads();

    // ---- 4: mask tiles -> HBM
    for (int i = tid; i < n_mask; i += FT_THREADS) {
        const int yy = i / mw, wx = i - yy * mw;
        const int y = ty0 + yy, word = (tx0 >> 5) + wx;
        if (y < H && word < p.ww) {
            const int64_t o = ((int64_t)f * H + y) * p.ww + word;
            p.mask_bits[o] = smask[i];
            if (p.marker_bits) p.marker_bits[o] = smask[n_mask + i];
        }
    }
}

// Tile geometry and the polarised threshold.  Decision bit = (d > t) != inverted, d = b - rint(mean).
//   white-on-dark: certain background  <=>  d <= t for every t in use  <=  mean >= b - t* - 0.49, t* = min t
//   dark-on-light: certain background  <=>  d >  t for every t in use  <=  mean <= b - T - 0.51,  T = max t; on complemented
//                  bytes p = 255 - b this reads  mean_p >= p + T + 0.51  =  p - (-T - 1) - 0.49
// both: candidate iff p > floor(L + t_q + 0.48), with 0.01 (0.02) left for the float32 rounding of the reference's mean.
static FusedGeom fused_geometry(const FrontParams &p)
{
    FusedGeom g{};
    g.tw = 256; g.th = 64;
    if (p.w <= 128) g.tw = 128;
    g.tiles_x = (p.w + g.tw - 1) / g.tw; g.tiles_y = (p.h + g.th - 1) / g.th;
    g.gp = g.tw + 24;
    while (g.gp % 8 != 0 || (g.gp / 8) % 2 == 0) g.gp += 4;
    g.rp = g.th + 8;
    while (g.rp % 4 != 0 || (g.rp / 4) % 2 == 0) g.rp += 4;
    const bool two = p.marker_bits != nullptr;
    if (!p.inverted) g.t_q = two ? (p.t_mask < p.t_marker ? p.t_mask : p.t_marker) : p.t_mask;
    else g.t_q = -(two ? (p.t_mask > p.t_marker ? p.t_mask : p.t_marker) : p.t_mask) - 1;
    const int grey_bytes = (g.th + 12) * g.gp, rq_bytes = (g.tw / 2) * g.rp;
    const int list_off = (rq_bytes + 15) & ~15, list_bytes = 2 * g.tw * g.th;
    int region_a = grey_bytes > list_off + list_bytes ? grey_bytes : list_off + list_bytes;
    region_a = (region_a + 15) & ~15;
    g.off_list = list_off;
    g.off_blur = region_a;
    g.off_mask = g.off_blur + (((g.th + 10) * g.gp + 15) & ~15);
    g.off_misc = g.off_mask + 2 * g.th * (g.tw / 32) * 4;
    g.smem_bytes = g.off_misc + 16;
    return g;
}

bool fused_frontend_supported(const FrontParams &p)
{
    return !p.scalar_thr && p.w % 4 == 0 && p.w >= 32 && (reinterpret_cast<uintptr_t>(p.frames) & 3) == 0 && p.frame_stride % 4 == 0;
}

// per device (the attribute is per device / context): called from ysmr_create
cudaError_t fused_frontend_init()
{
    cudaError_t e = cudaFuncSetAttribute(fused_front_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fused_front_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

cudaError_t launch_fused_frontend(const FrontParams &p, cudaStream_t st)
{
    const FusedGeom g = fused_geometry(p);
    const int64_t ctas = (int64_t)g.tiles_x * g.tiles_y * p.n_frames;
    if (ctas > 0x7fffffff || g.smem_bytes > 100 * 1024) return cudaErrorInvalidConfiguration;
    if (p.channels == 3) fused_front_kernel<3><<<(unsigned)ctas, FT_THREADS, g.smem_bytes, st>>>(p, g);
    else fused_front_kernel<1><<<(unsigned)ctas, FT_THREADS, g.smem_bytes, st>>>(p, g);
    return cudaGetLastError();
}

}  // namespace ysmr
