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
#include <stdlib.h>
#include <string.h>

#include "frontend.cuh"
#include "front_arith.cuh"

namespace ysmr {

constexpr int FT_THREADS = 256;

struct FusedGeom {
    int tw, th, ltw;            // tile size: tw = 1 << ltw (128 or 256), th % 4 == 0, tw * th <= 65536
    int tiles_x, tiles_y;
    int gp;                     // byte pitch of the grey / blurred tiles: >= tw + 24, gp % 8 == 0, (gp / 8) odd
    int rp;                     // byte pitch of the transposed row-pass bytes: >= th + 8, rp % 4 == 0, (rp / 4) odd
    int t_q;                    // polarised decision threshold (see fused_geometry)
    int off_blur, off_mask, off_list, off_misc;   // shared-memory offsets (bytes); grey tile and row-pass bytes at 0
    int list_cap;               // candidate list entries
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

// What the exact decision of one pixel needs (phase 3 and the overflow path of phase 2b).
struct RefineCtx {
    uint32_t s_blur;
    int gp, tw, ltw, mw, n_mask, tx0;
    int row_tail_from, col_tail_from, t_mask, t_marker;
    bool inv, two;
    uint32_t *smask;
};

// Exact decisions of tile pixel (xx, yy): OpenCV's float32 11x11 Gaussian (front_arith.cuh) around the pixel from the
// blurred tile, rint, both threshold compares; sets the pixel's bits in the shared mask tiles.
__device__ __noinline__ void refine_px(const RefineCtx &rc, int idx)
{
    const int yy = idx >> rc.ltw, xx = idx & (rc.tw - 1);
    const int x = rc.tx0 + xx;
    const bool row_tail = x >= rc.row_tail_from, col_tail = x >= rc.col_tail_from;
    const uint32_t a0 = rc.s_blur + yy * rc.gp + 12 + xx - 5;     // tile row yy + 5 - 5, byte column of x - 5
    const uint32_t al = a0 & ~3u;
    const uint32_t sel = 0x3210u + 0x1111u * (a0 & 3u);
    float r[11];
    int bc = 0;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
        const uint32_t ra = al + j * rc.gp;
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
    const bool m_mask = (d > rc.t_mask) != rc.inv, m_mark = (d > rc.t_marker) != rc.inv;
    const int wi = yy * rc.mw + (xx >> 5);
    const uint32_t bit = 1u << (xx & 31);
    if (m_mask) atomicOr(&rc.smask[wi], bit);
    if (m_mark && rc.two) atomicOr(&rc.smask[rc.n_mask + wi], bit);
}

extern __shared__ __align__(16) unsigned char ysmr_fused_smem[];

// gp of fused_geometry for a tile width (a compile-time constant inside the kernel: the tile width is a template parameter,
// so every division / multiplication by it or by the pitches derived from it folds)
__host__ __device__ constexpr int fused_gp(int tw)
{
    int gp = tw + 24;
    while (gp % 8 != 0 || (gp / 8) % 2 == 0) gp += 4;
    return gp;
}

// CS: the input pixels are read exactly once -- cache-streaming loads (ld.global.cs, evict-first in L2) keep them from
// pushing the masks, detection records and linker tables of the other kernels out of the 126 MB L2.
template <bool CS>
__device__ __forceinline__ uint32_t ld_px(const uint32_t *q) { return CS ? __ldcs(q) : __ldg(q); }

template <int C, int LTW, bool CS>
__global__ void __launch_bounds__(FT_THREADS, 3) fused_front_kernel(FrontParams p, FusedGeom g)
{
    const int tid = threadIdx.x, lane = tid & 31;
    // tile of this CTA: grid = (tiles_x, tiles_y, frames)
    const int tix = blockIdx.x, tiy = blockIdx.y, f = blockIdx.z;
    constexpr int tw = 1 << LTW, gp = fused_gp(tw);
    const int th = g.th, rp = g.rp;
    const int tx0 = tix * tw, ty0 = tiy * th;
    const int W = p.w, H = p.h;
    const uint8_t *frame = p.frames + (int64_t)f * p.frame_stride;

    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(ysmr_fused_smem);
    const uint32_t s_grey = s_base, s_rq = s_base, s_blur = s_base + g.off_blur, s_list = s_base + g.off_list;
    uint32_t *smask = reinterpret_cast<uint32_t *>(ysmr_fused_smem + g.off_mask);
    FusedMisc *misc = reinterpret_cast<FusedMisc *>(ysmr_fused_smem + g.off_misc);
    constexpr int lmw = LTW - 5, mw = 1 << lmw;                   // mask words per tile row
    const int n_mask = th * mw;

    {   // (n_mask is a multiple of 4 words: th % 4 == 0; the mask tiles are 16-byte aligned)
        uint4 *sm4 = reinterpret_cast<uint4 *>(smask);
        for (int i = tid; i < n_mask / 2; i += FT_THREADS) sm4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) { misc->tmin = 0xFFFFu; misc->tmax = 0u; misc->count = 0u; }

    // ---- 1a: grey tile.  Row r <-> virtual image row ty0 - 6 + r, byte c <-> virtual column tx0 - 12 + c.  The blur reads
    // one pixel beyond the image (REFLECT_101: column -1 is column 1, column W is column W-2, same for rows); everything
    // further out is never used, so out-of-image words are whole-word loads of the nearest image word with one byte moved.
    // A thread owns one pair of words (8 pixels) of every n-th row: column clamps and byte selectors are per-thread constants.
    {
        const int npair = (tw + 24) >> 3;                         // word pairs per row
        const int rstep = FT_THREADS / npair;
        const int n_rows = th + 12;
        const bool y_inside = ty0 - 6 >= 0 && ty0 + th + 6 <= H;
        const int r0 = tid / npair, cp = tid - r0 * npair;
        if (r0 < rstep) {
            const int vxa = tx0 - 12 + 8 * cp, vxb = vxa + 4;
            const int offa = min(max(vxa, 0), W - 4) * C, offb = min(max(vxb, 0), W - 4) * C;
            const uint32_t sela = vxa == -4 ? 0x1210u : (vxa == W ? 0x3212u : 0x3210u);
            const uint32_t selb = vxb == -4 ? 0x1210u : (vxb == W ? 0x3212u : 0x3210u);
            const int64_t rowstride = (int64_t)W * C;
            uint32_t so = s_grey + r0 * gp + 8 * cp;
            auto load = [&](int r, uint32_t (&raw)[2][C == 3 ? 3 : 1]) {
                int gy = ty0 - 6 + r;
                if (!y_inside) gy = reflect101(gy, H);
                const uint8_t *rowp = frame + gy * rowstride;
                const uint32_t *qa = reinterpret_cast<const uint32_t *>(rowp + offa), *qb = reinterpret_cast<const uint32_t *>(rowp + offb);
                raw[0][0] = ld_px<CS>(qa); raw[1][0] = ld_px<CS>(qb);
                if (C == 3) { raw[0][1] = ld_px<CS>(qa + 1); raw[0][2] = ld_px<CS>(qa + 2); raw[1][1] = ld_px<CS>(qb + 1); raw[1][2] = ld_px<CS>(qb + 2); }
            };
            auto store = [&](const uint32_t (&raw)[2][C == 3 ? 3 : 1]) {
                const uint32_t va = C == 3 ? grey4_of_bgr(raw[0][0], raw[0][1], raw[0][2]) : raw[0][0];
                const uint32_t vb = C == 3 ? grey4_of_bgr(raw[1][0], raw[1][1], raw[1][2]) : raw[1][0];
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(so), "r"(__byte_perm(va, va, sela)), "r"(__byte_perm(vb, vb, selb)));
                so += rstep * gp;
            };
            typedef uint32_t Raw[2][C == 3 ? 3 : 1];
            if (y_inside) {
                // interior tile (the common case): every row is an image row, the row pointer just advances, and rows
                // stream through a three-deep register pipeline (two rows of loads in flight while one is converted)
                const int gy0 = ty0 - 6 + r0;
                const uint8_t *pa = frame + gy0 * rowstride + offa;
                const int64_t dab = offb - offa, adv = rstep * rowstride;
                auto ld = [&](Raw &raw) {
                    const uint32_t *qa = reinterpret_cast<const uint32_t *>(pa), *qb = reinterpret_cast<const uint32_t *>(pa + dab);
                    raw[0][0] = ld_px<CS>(qa); raw[1][0] = ld_px<CS>(qb);
                    if (C == 3) { raw[0][1] = ld_px<CS>(qa + 1); raw[0][2] = ld_px<CS>(qa + 2); raw[1][1] = ld_px<CS>(qb + 1); raw[1][2] = ld_px<CS>(qb + 2); }
                    pa += adv;
                };
                const int cnt = (n_rows - r0 + rstep - 1) / rstep;       // rows of this thread (>= 3: th >= 16)
                Raw a, b, c;
                ld(a); ld(b);
                int k = 0;
                for (; k + 3 <= cnt - 2; k += 3) {
                    ld(c); store(a);
                    ld(a); store(b);
                    ld(b); store(c);
                }
                // drain: rows k .. cnt-1 (a, b hold rows k, k+1)
                if (k + 2 < cnt) { ld(c); store(a); if (k + 3 < cnt) { ld(a); store(b); store(c); store(a); } else { store(b); store(c); } }
                else { store(a); store(b); }
            } else {
                // software pipeline: the loads of the next three rows are issued before the current three are converted
                Raw x0, x1, x2, y0, y1, y2;
                int r = r0;
                const int step3 = 3 * rstep;
                auto load3 = [&](int rr, Raw &a, Raw &b, Raw &c) {
                    if (rr < n_rows) load(rr, a);
                    if (rr + rstep < n_rows) load(rr + rstep, b);
                    if (rr + 2 * rstep < n_rows) load(rr + 2 * rstep, c);
                };
                auto store3 = [&](int rr, const Raw &a, const Raw &b, const Raw &c) {
                    if (rr < n_rows) store(a);
                    if (rr + rstep < n_rows) store(b);
                    if (rr + 2 * rstep < n_rows) store(c);
                };
                load3(r, x0, x1, x2);
                for (; r < n_rows; r += 2 * step3) {
                    load3(r + step3, y0, y1, y2);
                    store3(r, x0, x1, x2);
                    load3(r + 2 * step3, x0, x1, x2);
                    store3(r + step3, y0, y1, y2);
                }
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
            const int yy = i >> LTW, xx = i & (tw - 1);
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
    const int sh = max(0, 24 - __clz(BH_SUM * range));            // smallest shift with (BH_SUM * range) >> sh <= 255
    const int acc_row = inv ? BH_SUM * tmax : -BH_SUM * tmin;     // start of the row chain: sum h (p - base_p) >= 0
    const int ka = inv ? (int)KAn : (int)KA, kb = inv ? (int)KBn : (int)KB, kc = inv ? (int)KCn : (int)KC;
    const int ka2 = inv ? (int)KA2n : (int)KA2, kb2 = inv ? (int)KB2n : (int)KB2, kc2 = inv ? (int)KC2n : (int)KC2;
    constexpr int nq2 = tw >> 2;                                     // row-pass bytes of pixel pair X live in column
                                                                  // X/2 (X even) or nq2 + X/2 (X odd) of the rq array

    // ---- 2a: row pass.  Task (q, G): pixels 8q .. 8q+7 of the four blurred rows ry = 4G .. 4G+3 (tile row ry + 1, image row
    // ty0 - 4 + ry) -> four pair results X = 4q .. 4q+3 per row, shifted to a byte, the four rows of a pair as one word of the
    // transposed array rq[col(X)][ry].  Lanes run along q: consecutive LDS.64.
    {
        constexpr int lq = LTW - 3, nq = 1 << lq; const int nG = (th + 8) >> 2;
        for (int t = tid; t < nq * nG; t += FT_THREADS) {
            const int q = t & (nq - 1), G = t >> lq;
            uint32_t a = s_blur + (4 * G + 1) * gp + 8 * q + 8;                   // relative words 2q-1 .. 2q+2
            int r0[4], r1[4], r2[4], r3[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint2 lo = lds64(a), hi = lds64(a + 8);
                a += gp;
                r0[k] = dp4a_us(lo.x, ka, acc_row); r0[k] = dp4a_us(lo.y, kb, r0[k]); r0[k] = dp4a_us(hi.x, kc, r0[k]);
                r1[k] = dp4a_us(lo.x, ka2, acc_row); r1[k] = dp4a_us(lo.y, kb2, r1[k]); r1[k] = dp4a_us(hi.x, kc2, r1[k]);
                r2[k] = dp4a_us(lo.y, ka, acc_row); r2[k] = dp4a_us(hi.x, kb, r2[k]); r2[k] = dp4a_us(hi.y, kc, r2[k]);
                r3[k] = dp4a_us(lo.y, ka2, acc_row); r3[k] = dp4a_us(hi.x, kb2, r3[k]); r3[k] = dp4a_us(hi.y, kc2, r3[k]);
            }
            // four results (< 2^16) -> four bytes (result >> sh < 256): two results per register as 16-bit fields, one shift
            // (the low bits of the upper field fall into byte 1, which is dropped), bytes 0 and 2 of both registers
            auto pack = [&](const int (&r)[4]) {
                const uint32_t lo = __byte_perm((uint32_t)r[0], (uint32_t)r[1], 0x5410) >> sh;
                const uint32_t hi = __byte_perm((uint32_t)r[2], (uint32_t)r[3], 0x5410) >> sh;
                return __byte_perm(lo, hi, 0x6420);
            };
            const uint32_t o0 = pack(r0), o1 = pack(r1), o2 = pack(r2), o3 = pack(r3);
            const uint32_t o = s_rq + (2 * q) * rp + 4 * G;
            sts32(o, o0); sts32(o + nq2 * rp, o1); sts32(o + rp, o2); sts32(o + (nq2 + 1) * rp, o3);
        }
    }
    __syncthreads();

    RefineCtx rc;
    rc.s_blur = s_blur; rc.gp = gp; rc.tw = tw; rc.ltw = LTW; rc.mw = mw; rc.n_mask = n_mask; rc.tx0 = tx0;
    rc.row_tail_from = p.row_tail_from; rc.col_tail_from = p.col_tail_from; rc.t_mask = p.t_mask; rc.t_marker = p.t_marker;
    rc.inv = inv; rc.two = p.marker_bits != nullptr; rc.smask = smask;
    const int list_cap = g.list_cap;
    auto push = [&](int idx) {
        const uint32_t k = atomicAdd(&misc->count, 1u);
        if ((int)k < list_cap) asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_list + 2 * k), "r"(idx));
        else refine_px(rc, idx);                                  // list full (a tile that is mostly foreground): decide right away
    };

    // ---- 2b: column pass + candidate test.  Task (j, m): the four pixels 4j .. 4j+3 (pairs X = 2j, 2j+1) of output rows
    // 4m .. 4m+3 (row pairs 2m, 2m+1).  With q = p - base_p (0 .. range) a pixel is certain background iff
    //     q <= T,  T = floor(t_q + 0.48 + L),  L = (column sum << sh) / 65536.
    // range <= 127 (any tile that is not saturated): byte-parallel test  q + A >= 128  <=>  q > T  with A = clamp(127 - T, 0, 128),
    // and A comes straight out of the dot-product chain run with negated weights.
    {
        constexpr int lj = LTW - 2, nj = 1 << lj; const int nm = th >> 2;
        const int ts = 16 - sh;
        const int c0 = (g.t_q * 65536 + 31457) >> sh;             // arithmetic shift: floor, i.e. towards "candidate"
        const bool swar = range <= 127;
        const int c1 = (128 << ts) - 1 - c0;
        const int sgn = inv ? -1 : 1;
        const uint32_t cq = inv ? (uint32_t)tmax * 0x01010101u : 0u - (uint32_t)tmin * 0x01010101u;
        for (int t = tid; t < nj * nm; t += FT_THREADS) {
            const int j = t & (nj - 1), m = t >> lj;
            if (tx0 + 4 * j >= W) continue;                       // (W % 4 == 0: a word is inside or outside)
            const uint32_t a = s_rq + j * rp + 4 * m;
            const uint32_t e0 = lds32(a), e1 = lds32(a + 4), e2 = lds32(a + 8);
            const uint32_t d0 = lds32(a + nq2 * rp), d1 = lds32(a + nq2 * rp + 4), d2 = lds32(a + nq2 * rp + 8);
            const uint32_t ba = s_blur + (4 * m + 5) * gp + 12 + 4 * j;
            const int y0 = ty0 + 4 * m;
            if (swar) {
                int aa0 = dp4a_us(e0, (int)KAn, c1); aa0 = dp4a_us(e1, (int)KBn, aa0); aa0 = dp4a_us(e2, (int)KCn, aa0);
                int ab0 = dp4a_us(e0, (int)KA2n, c1); ab0 = dp4a_us(e1, (int)KB2n, ab0); ab0 = dp4a_us(e2, (int)KC2n, ab0);
                int aa1 = dp4a_us(d0, (int)KAn, c1); aa1 = dp4a_us(d1, (int)KBn, aa1); aa1 = dp4a_us(d2, (int)KCn, aa1);
                int ab1 = dp4a_us(d0, (int)KA2n, c1); ab1 = dp4a_us(d1, (int)KB2n, ab1); ab1 = dp4a_us(d2, (int)KC2n, ab1);
                aa0 = __vimin_s32_relu(aa0 >> ts, 128); ab0 = __vimin_s32_relu(ab0 >> ts, 128);
                aa1 = __vimin_s32_relu(aa1 >> ts, 128); ab1 = __vimin_s32_relu(ab1 >> ts, 128);
                const uint32_t A_top = __byte_perm((uint32_t)aa0, (uint32_t)aa1, 0x4400);     // rows 4m, 4m+1
                const uint32_t A_bot = __byte_perm((uint32_t)ab0, (uint32_t)ab1, 0x4400);     // rows 4m+2, 4m+3
                // z = q + A per byte; bit 7 set <=> candidate.  cq is folded into A; all four rows are tested with one branch.
                const uint32_t ct = cq + A_top, cb = cq + A_bot;
                const uint32_t z0 = (uint32_t)sgn * lds32(ba) + ct, z1 = (uint32_t)sgn * lds32(ba + gp) + ct;
                const uint32_t z2 = (uint32_t)sgn * lds32(ba + 2 * gp) + cb, z3 = (uint32_t)sgn * lds32(ba + 3 * gp) + cb;
                if ((z0 | z1 | z2 | z3) & 0x80808080u) {
                    // one slot reservation for all candidates of the task, then plain stores
                    uint32_t zz[4] = {z0 & 0x80808080u, z1 & 0x80808080u, z2 & 0x80808080u, z3 & 0x80808080u};
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (y0 + k >= H) zz[k] = 0;
                    const int n = __popc(zz[0]) + __popc(zz[1]) + __popc(zz[2]) + __popc(zz[3]);
                    uint32_t slot = n ? atomicAdd(&misc->count, (uint32_t)n) : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t z = zz[k];
                        while (z) {
                            const int byte = (__ffs(z) - 1) >> 3;
                            z &= z - 1;
                            const int idx = (4 * m + k) * tw + 4 * j + byte;
                            if ((int)slot < list_cap) asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_list + 2 * slot), "r"(idx));
                            else refine_px(rc, idx);
                            ++slot;
                        }
                    }
                }
            } else {
                // saturated tile (range of the blurred tile > 127): scalar compares
                int la0 = dp4a_us(e0, (int)KA, c0); la0 = dp4a_us(e1, (int)KB, la0); la0 = dp4a_us(e2, (int)KC, la0);
                int lb0 = dp4a_us(e0, (int)KA2, c0); lb0 = dp4a_us(e1, (int)KB2, lb0); lb0 = dp4a_us(e2, (int)KC2, lb0);
                int la1 = dp4a_us(d0, (int)KA, c0); la1 = dp4a_us(d1, (int)KB, la1); la1 = dp4a_us(d2, (int)KC, la1);
                int lb1 = dp4a_us(d0, (int)KA2, c0); lb1 = dp4a_us(d1, (int)KB2, lb1); lb1 = dp4a_us(d2, (int)KC2, lb1);
                la0 >>= ts; lb0 >>= ts; la1 >>= ts; lb1 >>= ts;
                for (int k = 0; k < 4; ++k) {
                    if (y0 + k >= H) break;
                    const uint32_t q4 = (uint32_t)sgn * lds32(ba + k * gp) + cq;
                    const int t0 = k < 2 ? la0 : lb0, t1 = k < 2 ? la1 : lb1;
                    for (int b = 0; b < 4; ++b)
                        if ((int)((q4 >> (8 * b)) & 0xFFu) > (b < 2 ? t0 : t1)) push((4 * m + k) * tw + 4 * j + b);
                }
            }
        }
    }
    __syncthreads();

    // ---- 3: exact decisions of the candidates (OpenCV's float32 arithmetic, front_arith.cuh)
    {
        const int n_cand = min((int)misc->count, list_cap);
        // candidate i goes to warp i % 8, lane (i / 8) % 32: a tile with a handful of candidates keeps all warps (and all
        // four schedulers) busy instead of one
        const int w = tid >> 5;
        for (int i = w + 8 * lane; i < n_cand; i += FT_THREADS) {
            uint32_t idx;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(idx) : "r"(s_list + 2 * i));
            refine_px(rc, (int)idx);
        }
    }
    __syncthreads();

    // ---- 4: mask tiles -> HBM
    for (int i = tid; i < n_mask; i += FT_THREADS) {
        const int yy = i >> lmw, wx = i & (mw - 1);
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
    g.tw = p.w <= 128 ? 128 : 256; g.ltw = p.w <= 128 ? 7 : 8;
    // tile height 104: the tallest tile with three CTAs per SM (71 KB of shared memory each); measured best of 56 .. 120 on
    // 1228 x 922 (taller tiles amortise the 12-row halo and the per-phase set-up, shorter ones hide more latency)
    g.th = 104;
    if (p.h < g.th) g.th = (p.h + 3) & ~3;
    g.tiles_x = (p.w + g.tw - 1) / g.tw; g.tiles_y = (p.h + g.th - 1) / g.th;
    g.gp = g.tw + 24;
    while (g.gp % 8 != 0 || (g.gp / 8) % 2 == 0) g.gp += 4;
    g.rp = g.th + 8;
    while (g.rp % 4 != 0 || (g.rp / 4) % 2 == 0) g.rp += 4;
    const bool two = p.marker_bits != nullptr;
    if (!p.inverted) g.t_q = two ? (p.t_mask < p.t_marker ? p.t_mask : p.t_marker) : p.t_mask;
    else g.t_q = -(two ? (p.t_mask > p.t_marker ? p.t_mask : p.t_marker) : p.t_mask) - 1;
    // shared memory: [grey tile | later: row-pass bytes, then the candidate list] [blurred tile] [mask tiles] [misc]
    const int grey_bytes = (g.th + 12) * g.gp, rq_bytes = (g.tw / 2) * g.rp;
    g.off_list = (rq_bytes + 15) & ~15;
    int region_a = (grey_bytes + 15) & ~15;
    if (region_a < g.off_list + 4096) region_a = g.off_list + 4096;
    g.list_cap = (region_a - g.off_list) / 2;
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
template <int C, int LTW, bool CS>
static cudaError_t fused_attr()
{
    return cudaFuncSetAttribute(fused_front_kernel<C, LTW, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}
cudaError_t fused_frontend_init()
{
    cudaError_t e = fused_attr<1, 7, false>();
    if (e == cudaSuccess) e = fused_attr<1, 8, false>();
    if (e == cudaSuccess) e = fused_attr<3, 7, false>();
    if (e == cudaSuccess) e = fused_attr<3, 8, false>();
    if (e == cudaSuccess) e = fused_attr<1, 7, true>();
    if (e == cudaSuccess) e = fused_attr<1, 8, true>();
    if (e == cudaSuccess) e = fused_attr<3, 7, true>();
    if (e == cudaSuccess) e = fused_attr<3, 8, true>();
    return e;
}

template <bool CS>
static void launch_fused_cs(const FrontParams &p, const FusedGeom &g, dim3 grid, cudaStream_t st)
{
    if (p.channels == 3) {
        if (g.ltw == 8) fused_front_kernel<3, 8, CS><<<grid, FT_THREADS, g.smem_bytes, st>>>(p, g);
        else fused_front_kernel<3, 7, CS><<<grid, FT_THREADS, g.smem_bytes, st>>>(p, g);
    } else {
        if (g.ltw == 8) fused_front_kernel<1, 8, CS><<<grid, FT_THREADS, g.smem_bytes, st>>>(p, g);
        else fused_front_kernel<1, 7, CS><<<grid, FT_THREADS, g.smem_bytes, st>>>(p, g);
    }
}

cudaError_t launch_fused_frontend(const FrontParams &p, cudaStream_t st)
{
    const FusedGeom g = fused_geometry(p);
    if (g.tiles_y > 65535 || p.n_frames > 65535 || g.smem_bytes > 100 * 1024 || g.gp != fused_gp(g.tw)) return cudaErrorInvalidConfiguration;
    const dim3 grid((unsigned)g.tiles_x, (unsigned)g.tiles_y, (unsigned)p.n_frames);
    static const bool plain_loads = [] { const char *e = getenv("YSMR_FUSED_LOADS"); return e && strcmp(e, "ldg") == 0; }();
    if (plain_loads) launch_fused_cs<false>(p, g, grid, st);     // (measurement: the A/B of the streaming loads)
    else launch_fused_cs<true>(p, g, grid, st);
    return cudaGetLastError();
}

}  // namespace ysmr
