// K1 -- fused front-end: BGR->grey, 3x3 binomial blur, 11x11 float32 Gaussian mean, threshold compare(s), bit-packing.
//
// Replaces, per frame, /root/reference/ysmr/track_eval.py:180 (cv2.cvtColor), :182 (cv2.GaussianBlur 3x3),
// :189-197 and :200-208 (two cv2.adaptiveThreshold calls -- the reference computes the identical Gaussian mean image
// twice; here it is computed once and compared against both constants) and, in the mean/std mode, :248-253
// (cv2.threshold).  Arithmetic is bit-exact with OpenCV 4.13 (oracle/c_stages.c): integer luma and
// blur, float32 Gaussian with OpenCV's FMA placement including its scalar-tail columns, rint half-to-even.
//
// Output: 1 bit per pixel masks (bit x&31 of word [y][x>>5]) so that the labelling kernel reads 1/8 of a byte image.
//
// This file holds two implementations that must agree bit for bit:
//   frontend_tile_kernel   -- straightforward shared-memory tiles, can also dump every intermediate stage (debug);
//   K1a / K1b / K1c        -- the production kernels (blur pre-pass, Gaussian + decisions, mask packing), see below.
#include "frontend.cuh"
#include "front_arith.cuh"

namespace ysmr {


// ---------------------------------------------------------------------------------------------------------------------
// Tile kernel (debug / reference implementation on the device)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TILE_W = 64, TILE_H = 32, HALO = 6;
constexpr int G_W = TILE_W + 2 * HALO, G_H = TILE_H + 2 * HALO;          // grey tile 76 x 44
constexpr int B_W = TILE_W + 10, B_H = TILE_H + 10;                      // blurred tile 74 x 42

template <int C>
__global__ void __launch_bounds__(256) frontend_tile_kernel(FrontParams p)
{
    __shared__ uint8_t s_g[G_H][G_W + 4];
    __shared__ float s_b[B_H][B_W + 2];
    __shared__ float s_r[B_H][TILE_W];
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TILE_W, ty0 = blockIdx.y * TILE_H, f = blockIdx.z;
    const uint8_t *frame = p.frames + (int64_t)f * p.frame_stride;
    const int64_t plane = (int64_t)p.h * p.w;

    for (int i = tid; i < G_H * G_W; i += 256) {
        const int ly = i / G_W, lx = i - ly * G_W;
        const int gy = reflect101(ty0 - HALO + ly, p.h), gx = reflect101(tx0 - HALO + lx, p.w);
        const uint8_t *px = frame + ((int64_t)gy * p.w + gx) * C;
        s_g[ly][lx] = C == 3 ? (uint8_t)luma(px[0], px[1], px[2]) : px[0];
    }
    __syncthreads();
    for (int i = tid; i < B_H * B_W; i += 256) {
        const int ly = i / B_W, lx = i - ly * B_W;
        const int gy = ty0 - 5 + ly, gx = tx0 - 5 + lx;
        const int cy = clampi(gy, p.h) - (ty0 - HALO), cx = clampi(gx, p.w) - (tx0 - HALO);   // BORDER_REPLICATE of the blur
        const int s = s_g[cy - 1][cx - 1] + 2 * s_g[cy - 1][cx] + s_g[cy - 1][cx + 1] +
                      2 * s_g[cy][cx - 1] + 4 * s_g[cy][cx] + 2 * s_g[cy][cx + 1] +
                      s_g[cy + 1][cx - 1] + 2 * s_g[cy + 1][cx] + s_g[cy + 1][cx + 1];
        const int b = (s + 8) >> 4;
        s_b[ly][lx] = (float)b;
        if (gy >= ty0 && gy < ty0 + TILE_H && gy < p.h && gx >= tx0 && gx < tx0 + TILE_W && gx < p.w) {
            if (p.dbg_grey) p.dbg_grey[f * plane + (int64_t)gy * p.w + gx] = s_g[cy][cx];
            if (p.dbg_blurred) p.dbg_blurred[f * plane + (int64_t)gy * p.w + gx] = (uint8_t)b;
        }
    }
    __syncthreads();
    for (int i = tid; i < B_H * TILE_W; i += 256) {
        const int ly = i / TILE_W, ox = i - ly * TILE_W;
        const bool tail = tx0 + ox >= p.row_tail_from;
        float acc = __fmul_rn(p.k[0], s_b[ly][ox]);
#pragma unroll
        for (int t = 1; t < 11; ++t) {
            const float v = s_b[ly][ox + t];
            if (!tail || t >= 9) acc = __fmaf_rn(p.k[t], v, acc);
            else acc = __fadd_rn(acc, __fmul_rn(p.k[t], v));
        }
        s_r[ly][ox] = acc;
    }
    __syncthreads();
    const int ox = tid & 63;
    const int x = tx0 + ox;
    const bool tail = x >= p.col_tail_from;
    const int thr_scalar = p.scalar_thr ? p.scalar_thr[f] : 0;
    for (int oy = tid >> 6; oy < TILE_H; oy += 4) {
        const int y = ty0 + oy;
        float acc = __fmul_rn(p.k[5], s_r[oy + 5][ox]);
#pragma unroll
        for (int j = 1; j <= 5; ++j) {
            const float s = __fadd_rn(s_r[oy + 5 - j][ox], s_r[oy + 5 + j][ox]);
            if (!tail) acc = __fmaf_rn(p.k[5 + j], s, acc);
            else acc = __fadd_rn(acc, __fmul_rn(p.k[5 + j], s));
        }
        int mean = __float2int_rn(acc);
        mean = mean < 0 ? 0 : (mean > 255 ? 255 : mean);
        const int b = (int)s_b[oy + 5][ox + 5];
        const bool valid = x < p.w && y < p.h;
        bool m_mask, m_mark;
        if (p.scalar_thr) {
            m_mask = (b > thr_scalar) != (p.inverted != 0);
            m_mark = false;
        } else {
            const int d = b - mean;
            m_mask = (d > p.t_mask) != (p.inverted != 0);
            m_mark = (d > p.t_marker) != (p.inverted != 0);
        }
        const uint32_t w_mask = __ballot_sync(0xffffffffu, valid && m_mask);
        const uint32_t w_mark = __ballot_sync(0xffffffffu, valid && m_mark);
        if (valid && p.dbg_mean) p.dbg_mean[f * plane + (int64_t)y * p.w + x] = (uint8_t)mean;
        const int word = x >> 5;
        if ((tid & 31) == 0 && y < p.h && word < p.ww) {
            const int64_t o = ((int64_t)f * p.h + y) * p.ww + word;
            p.mask_bits[o] = w_mask;
            if (p.marker_bits) p.marker_bits[o] = w_mark;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Shared pieces of the production kernels
// ---------------------------------------------------------------------------------------------------------------------

// =====================================================================================================================
// Production front-end, generation 3: three kernels
//
//   K1a blur_prepass_kernel   BGR/grey -> luma -> 3x3 binomial blur, written as a PADDED u8 plane (replicated margins), so
//                             that the Gaussian kernel below has no border logic at all.  Integer work, HBM-bound.
//   K1b gauss_decide_kernel   11x11 float32 Gaussian mean (OpenCV's operation order), rint, the two threshold decisions.
//                             Everything after the u8->float conversion runs on the FP32 pipe (measured on B200: FFMA/FADD
//                             issue 1/clk per SM sub-partition, every integer ALU op 1 per 2 clk, SHFL 1 per 4 clk --
//                             scripts/ubench.cu), including the decisions (add.sat) and the packing of a lane's eight
//                             decisions into one byte (FFMA chain onto 2^23).  Output: one "decision byte" per lane and row.
//   K1c pack_masks_kernel     decision bytes -> the two standard bit masks (bit x&31 of word x>>5), polarity and width applied.
//
// Blurred plane of one frame: rows -5 .. h+4, `pitch` bytes each; pixel (x, y) at (y + 5) * pitch + 8 + x; columns -8..-1
// and w..w+7 and the 5 rows above / below hold BORDER_REPLICATE copies (what cv2.adaptiveThreshold's Gaussian sees).
// =====================================================================================================================
constexpr int PB_WARPS = 4;
#ifndef PB_MIN_CTAS
#define PB_MIN_CTAS 10
#endif
constexpr int PB_COLS = 256;                      // columns per warp of K1a: 8 adjacent pixels per lane


// One row of a lane before the blur: grey bytes of its 8 pixels (w0, w1) and the words holding the two side pixels.
struct RowBytes {
    uint32_t w0, w1, wl, wr;
};

// Per-lane constants (FAST path: w % 4 == 0 and 4-byte aligned frames).  Every load is a whole 32-bit word at an address that
// is always inside the row, and REFLECT_101 at the image border as well as the position of the side pixels inside their
// words are folded into four byte-permute selectors, so the row loop has no branches and no selects.
struct LaneGeom {
    int hi_off;                  // word offset of pixels 4..7 (0 when the lane only has 4 valid pixels: w % 8 == 4, last lane)
    int l_off, r_off;            // word offsets of the words holding pixel gx-1 / gx+8 (0 at the image border)
    uint32_t sel_a0, sel_b0, sel_b1;     // (g-1,g0,g1,g2) from (wl,w0) ; (g1,g2,g3,g4) from (w0,w1) ; (g5,g6,g7,g8) from (w1,wr)
};

template <int C>
__device__ __forceinline__ LaneGeom lane_geom(int w, int gx)
{
    LaneGeom g;
    const bool half = w - gx < 8, left_edge = gx == 0, right_edge = !half && gx + 8 >= w;
    g.hi_off = half ? 0 : C;
    g.l_off = left_edge ? 0 : -1;
    g.r_off = (half || right_edge) ? 0 : 2 * C;
    // the left pixel is byte 3 of wl (C == 1: last byte of the previous word) or byte 2 (C == 3: the luma sum); at the left
    // image border column -1 is column 1 (REFLECT_101) = byte 1 of w0
    g.sel_a0 = left_edge ? 0x6545u : (C == 1 ? 0x6543u : 0x6542u);
    g.sel_b0 = half ? 0x2321u : 0x4321u;                           // w % 8 == 4, last lane: column w is column w-2
    // the right pixel is byte 0 of wr (C == 1) or byte 2 of the luma sum (C == 3); at the right border column w is w-2
    g.sel_b1 = right_edge ? 0x2321u : (C == 1 ? 0x4321u : 0x6321u);
    return g;
}

// Raw words of one row of a lane.  Loading and converting are separate steps so that the loads issued in one iteration are
// first touched in the next: a whole iteration of latency hiding per warp (the kernel is bound by bytes in flight).
template <int C>
struct RawRow {
    uint32_t w[C == 1 ? 4 : 8];                            // C == 1: w0, w1, wl, wr ; C == 3: 3 + 3 BGR words, left, right
};

template <int C>
__device__ __forceinline__ RawRow<C> fetch_row_fast(const uint8_t *rowp, const LaneGeom &lg)
{
    const uint32_t *q = reinterpret_cast<const uint32_t *>(rowp);
    RawRow<C> r;
    if (C == 1) {
        r.w[0] = __ldg(q); r.w[1] = __ldg(q + lg.hi_off); r.w[2] = __ldg(q + lg.l_off); r.w[3] = __ldg(q + lg.r_off);
    } else {
        const uint32_t *qh = q + 3 * lg.hi_off / C;        // hi_off is 0 or C -> 0 or 3 words
        r.w[0] = __ldg(q); r.w[1] = __ldg(q + 1); r.w[2] = __ldg(q + 2);
        r.w[3] = __ldg(qh); r.w[4] = __ldg(qh + 1); r.w[5] = __ldg(qh + 2);
        r.w[6] = __ldg(q + lg.l_off); r.w[7] = __ldg(q + 3 * lg.r_off / C);
    }
    return r;
}

template <int C>
__device__ __forceinline__ RowBytes bytes_of_raw(const RawRow<C> &x)
{
    RowBytes r;
    if (C == 1) {
        r.w0 = x.w[0]; r.w1 = x.w[1]; r.wl = x.w[2]; r.wr = x.w[3];
    } else {
        r.w0 = grey4_of_bgr(x.w[0], x.w[1], x.w[2]);
        r.w1 = grey4_of_bgr(x.w[3], x.w[4], x.w[5]);
        r.wl = lsum_b123(x.w[6]); r.wr = lsum_b012(x.w[7]);   // grey value in byte 2
    }
    return r;
}

// generic path: per-pixel loads with REFLECT_101 (any width / alignment, slow); same word layout as the FAST path
template <int C>
__device__ __forceinline__ RowBytes load_row_generic(const uint8_t *frame, int w, int y, int gx)
{
    uint32_t v[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) v[k] = grey_px<C>(frame, w, y, reflect101(gx - 1 + k, w));
    RowBytes r;
    r.w0 = v[1] | (v[2] << 8) | (v[3] << 16) | (v[4] << 24);
    r.w1 = v[5] | (v[6] << 8) | (v[7] << 16) | (v[8] << 24);
    r.wl = C == 1 ? v[0] << 24 : v[0] << 16;
    r.wr = C == 1 ? v[9] : v[9] << 16;
    return r;
}

// Horizontal 1-2-1 of one row as byte dot products: four packed pairs (h0,h1) .. (h6,h7), each value + 2 so that the
// vertical 1-2-1 (weights sum to 4) carries the rounding constant 8 of (sum + 8) >> 4.
struct HRow {
    uint32_t h01, h23, h45, h67;
};

__device__ __forceinline__ HRow hblur(const RowBytes &r, const LaneGeom &lg)
{
    const uint32_t a0 = __byte_perm(r.wl, r.w0, lg.sel_a0);       // (g-1, g0, g1, g2)
    const uint32_t b0 = __byte_perm(r.w0, r.w1, lg.sel_b0);       // (g1, g2, g3, g4)
    const uint32_t a1 = __byte_perm(r.w0, r.w1, 0x6543);          // (g3, g4, g5, g6)
    const uint32_t b1 = __byte_perm(r.w1, r.wr, lg.sel_b1);       // (g5, g6, g7, g8)
    constexpr uint32_t LO = 0x00010201u, HI = 0x01020100u;        // weights (1,2,1,0) and (0,1,2,1)
    HRow h;
    h.h01 = __dp4a(a0, LO, 2u) + (__dp4a(a0, HI, 2u) << 16);
    h.h23 = __dp4a(b0, LO, 2u) + (__dp4a(b0, HI, 2u) << 16);
    h.h45 = __dp4a(a1, LO, 2u) + (__dp4a(a1, HI, 2u) << 16);
    h.h67 = __dp4a(b1, LO, 2u) + (__dp4a(b1, HI, 2u) << 16);
    return h;
}

__device__ __forceinline__ HRow add_rows(const HRow &a, const HRow &b)
{
    HRow s;
    s.h01 = a.h01 + b.h01; s.h23 = a.h23 + b.h23; s.h45 = a.h45 + b.h45; s.h67 = a.h67 + b.h67;
    return s;
}

__device__ __forceinline__ uint2 blur_out(const HRow &v)          // v = h(y-1) + 2 h(y) + h(y+1) + 8 per 16-bit lane
{
    return make_uint2(__byte_perm(v.h01 >> 4, v.h23 >> 4, 0x6420), __byte_perm(v.h45 >> 4, v.h67 >> 4, 0x6420));
}

template <int C, bool FAST>
__global__ void __launch_bounds__(PB_WARPS * 32, PB_MIN_CTAS) blur_prepass_kernel(FrontParams p, int n_strips, int n_chunks, int rows_per_chunk)
{
    const int lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * PB_WARPS + (threadIdx.x >> 5);
    const int64_t per_frame = (int64_t)n_strips * n_chunks;
    if (task >= per_frame * p.n_frames) return;
    const int f = (int)(task / per_frame);
    const int r = (int)(task - (int64_t)f * per_frame);
    const int strip = r % n_strips, chunk = r / n_strips;
    const int y0 = chunk * rows_per_chunk, y1 = min(p.h, y0 + rows_per_chunk);
    const int w = p.w, h = p.h;
    const uint8_t *frame = p.frames + (int64_t)f * p.frame_stride;
    const int gx = strip * PB_COLS + 8 * lane;
    if (gx >= w) return;                                          // no cross-lane traffic in this kernel
    uint8_t *dst = p.plane + (int64_t)f * p.plane_stride + (int64_t)(y0 + 5) * p.pitch + 8 + gx;
    const int pitch = p.pitch;
    const int64_t stride = (int64_t)w * C;
    const uint8_t *col = frame + (int64_t)gx * C;                 // this lane's column in row 0
    LaneGeom lg = lane_geom<C>(w, gx);
    if (!FAST) { lg.sel_a0 = C == 1 ? 0x6543u : 0x6542u; lg.sel_b0 = 0x4321u; lg.sel_b1 = C == 1 ? 0x4321u : 0x6321u; }   // borders done by the loads
    // horizontal pass first (on bytes), then the vertical 1-2-1 slides through registers as S(y-1) = h(y-1) + h(y) and h(y)
    if (FAST) {
        auto fetch = [&](int y) { return fetch_row_fast<C>(col + y * stride, lg); };      // y already inside [0, h)
        auto conv = [&](const RawRow<C> &x) { return hblur(bytes_of_raw<C>(x), lg); };
        HRow b = conv(fetch(y0));
        HRow s_prev = add_rows(conv(fetch(reflect101(y0 - 1, h))), b);
        RawRow<C> raw = fetch(reflect101(y0 + 1, h));              // row y+1 of the first iteration, in flight
#pragma unroll 2
        for (int y = y0; y < y1; ++y) {
            int yn = y + 2; yn = yn >= h ? 2 * h - 2 - yn : yn;   // REFLECT_101 below the image (last chunk only)
            RawRow<C> ahead = raw;
            if (y + 1 < y1) ahead = fetch(yn);                    // issue the loads of row y+2 ...
            const HRow c = conv(raw);                             // ... and only now touch row y+1, loaded one iteration ago
            raw = ahead;
            const HRow s_cur = add_rows(b, c);                    // S(y) = h(y) + h(y+1)
            *reinterpret_cast<uint2 *>(dst) = blur_out(add_rows(s_prev, s_cur));
            s_prev = s_cur; b = c;
            dst += pitch;
        }
    } else {
        auto row = [&](int y) { return hblur(load_row_generic<C>(frame, w, y, gx), lg); };
        HRow b = row(y0);
        HRow s_prev = add_rows(row(reflect101(y0 - 1, h)), b);
        for (int y = y0; y < y1; ++y) {
            const HRow c = row(reflect101(y + 1, h));
            const HRow s_cur = add_rows(b, c);
            *reinterpret_cast<uint2 *>(dst) = blur_out(add_rows(s_prev, s_cur));
            s_prev = s_cur; b = c;
            dst += pitch;
        }
    }
}

// Replicated margins of the blurred planes (BORDER_REPLICATE of cv2.adaptiveThreshold's Gaussian): columns -8..-1 and
// w..w+7 of every image row, and the 5 rows above / below the image.  One work item per (frame, image row) plus one per
// (frame, margin row, 32-bit word).
__global__ void __launch_bounds__(256) plane_margins_kernel(FrontParams p)
{
    const int wpr = p.pitch / 4;                                  // words per plane row
    const int64_t per_frame = (int64_t)p.h + 10 * (int64_t)wpr;
    const int64_t total = per_frame * p.n_frames;
    const bool w4 = (p.w & 3) == 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i / per_frame);
        const int k = (int)(i - (int64_t)f * per_frame);
        uint8_t *plane = p.plane + (int64_t)f * p.plane_stride;
        if (k < p.h) {
            uint8_t *row = plane + (k + 5) * (int64_t)p.pitch + 8;
            uint32_t *row32 = reinterpret_cast<uint32_t *>(row);
            const uint32_t l = (row32[0] & 0xFFu) * 0x01010101u;
            row32[-2] = l; row32[-1] = l;
            if (w4) {
                const uint32_t r = (row32[p.w / 4 - 1] >> 24) * 0x01010101u;
                row32[p.w / 4] = r; row32[p.w / 4 + 1] = r;
            } else {
                const uint8_t r = row[p.w - 1];
                for (int x = p.w; x < p.w + 8; ++x) row[x] = r;
            }
        } else {
            const int kk = k - p.h;
            const int mr = kk / wpr, word = kk - mr * wpr;
            const int dst_row = mr < 5 ? mr : p.h + 5 + (mr - 5);              // plane row index (image row + 5)
            const int src_row = mr < 5 ? 5 : p.h + 4;
            const uint8_t *src = plane + (int64_t)src_row * p.pitch + 8;
            const int x0 = 4 * word - 8;
            uint32_t v;
            if (x0 >= 0 && x0 + 3 < p.w) {
                v = *reinterpret_cast<const uint32_t *>(src + x0);             // interior word: copy as is
            } else {
                v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    int x = x0 + b;
                    x = x < 0 ? 0 : (x >= p.w ? p.w - 1 : x);
                    v |= (uint32_t)src[x] << (8 * b);
                }
            }
            reinterpret_cast<uint32_t *>(plane + (int64_t)dst_row * p.pitch)[word] = v;
        }
    }
}

// ---- K1b ------------------------------------------------------------------------------------------------------------
// Work item = (frame, 128-column strip, row chunk), one per WARP; lane l owns columns x0 = xs + 4l .. +3.  Step s handles
// blurred row y0 - 5 + s: LDG.32 -> four floats -> ring slot s & 7 of the warp's shared row buffer; 5 x LDS.128 give the 14
// taps of the lane's four row-pass results, which enter an 11-row register window; the column pass around the row 5 steps
// back gives the float mean of output row y0 + s - 10.  Decisions, all exact in float:
//     R = mean + 1.5*2^23                      (rint, half to even, as an integer-valued float)
//     d = sat((b + 1.5*2^23 - t) - R)          (1.0f iff b - rint(mean) > t ; both operands are integers < 2^24)
// and the lane's byte = sum d_mask[k] 2^k + d_marker[k] 2^(4+k) accumulated onto 2^23 so that it is the low mantissa byte.
// The 2 x 8 halo columns of a strip are converted by all 32 lanes at once for 8 rows every 8 steps.
#ifndef GD_WARPS_N
#define GD_WARPS_N 4
#endif
#ifndef GD_MIN_CTAS
#define GD_MIN_CTAS 4
#endif
constexpr int GD_WARPS = GD_WARPS_N;
constexpr int GD_RING = 8;
constexpr int GD_ROWF = 144;                      // floats per ring slot: 8 halo | 128 | 8 halo
constexpr float GD_MAGIC = 12582912.0f;           // 1.5 * 2^23

__device__ __forceinline__ float4 u8x4_to_float4(uint32_t u)
{
    // 0x4B000000 | byte is the float 2^23 + byte
    const float2 neg = make_float2(-8388608.0f, -8388608.0f);
    const float2 a = fadd2(make_float2(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7540)), __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7541))), neg);
    const float2 b = fadd2(make_float2(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7542)), __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7543))), neg);
    return make_float4(a.x, a.y, b.x, b.y);
}

// One step of the register window: file this step's row-pass results in slot J and run the column pass centred 5 steps
// back.  Only this part depends on J (static register indices), so only this part is replicated 11 times behind the switch
// of the step loop -- the replicated code must stay small: with the whole step body inside the switch the kernel was bound by
// instruction-cache misses (ncu: stall_no_instruction 1.9 warps per issue, icc hit rate 73 %).
template <int J, bool TAIL>
__device__ __forceinline__ void gd_window(float2 (&win)[11][2], const float (&r)[4], float (&m)[4], bool col_tail)
{
    constexpr int c = (J + 6) % 11;               // (J - 5) mod 11: window slot of the centre row
    win[J][0] = make_float2(r[0], r[1]); win[J][1] = make_float2(r[2], r[3]);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float2 v = gauss_col2<TAIL>(win[c][k], win[(c + 10) % 11][k], win[(c + 1) % 11][k], win[(c + 9) % 11][k], win[(c + 2) % 11][k],
                                          win[(c + 8) % 11][k], win[(c + 3) % 11][k], win[(c + 7) % 11][k], win[(c + 4) % 11][k],
                                          win[(c + 6) % 11][k], win[(c + 5) % 11][k], col_tail);
        m[2 * k] = v.x; m[2 * k + 1] = v.y;
    }
}

// lane's decision byte from the four float means and the four blurred values of the output row.  Packed where the operation
// exists in packed form (FADD2 / FFMA2 halve the issue slots; sub.sat has no packed form):
//     R = mean + 1.5*2^23 ; e = b - R (exact: integers below 2^24) ; d = sat(e + (1.5*2^23 - t)) ; byte = sum d_k 2^k
__device__ __forceinline__ uint32_t gd_decide(const float (&m)[4], const float4 bf, float c_mask, float c_mark)
{
    // two steps on purpose: R rounds the mean to an integer first, then the difference is exact
    const float2 r01 = fadd2(make_float2(m[0], m[1]), make_float2(GD_MAGIC, GD_MAGIC));
    const float2 r23 = fadd2(make_float2(m[2], m[3]), make_float2(GD_MAGIC, GD_MAGIC));
    const float2 e01 = fadd2(make_float2(bf.x, bf.y), make_float2(-r01.x, -r01.y));
    const float2 e23 = fadd2(make_float2(bf.z, bf.w), make_float2(-r23.x, -r23.y));
    float2 da01, da23, db01, db23;
    asm("add.sat.f32 %0, %1, %2;" : "=f"(da01.x) : "f"(e01.x), "f"(c_mask));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(da01.y) : "f"(e01.y), "f"(c_mask));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(da23.x) : "f"(e23.x), "f"(c_mask));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(da23.y) : "f"(e23.y), "f"(c_mask));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(db01.x) : "f"(e01.x), "f"(c_mark));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(db01.y) : "f"(e01.y), "f"(c_mark));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(db23.x) : "f"(e23.x), "f"(c_mark));
    asm("add.sat.f32 %0, %1, %2;" : "=f"(db23.y) : "f"(e23.y), "f"(c_mark));
    // even pixels accumulate in .x, odd pixels in .y with the same (broadcast) weights; byte = x + 2 y on top of 2^23
    float2 acc = fadd2(da01, make_float2(8388608.0f, 0.0f));
    acc = ffma2(da23, make_float2(4.f, 4.f), acc);
    acc = ffma2(db01, make_float2(16.f, 16.f), acc);
    acc = ffma2(db23, make_float2(64.f, 64.f), acc);
    return __float_as_uint(__fmaf_rn(acc.y, 2.0f, acc.x));
}

// TAIL = false: strips without scalar-tail columns (every column fused).  TAIL = true: the strip(s) that contain OpenCV's
// scalar-tail columns (the last w % 8 columns of the column filter and the last w % 4 of the row filter are not
// FMA-contracted); launched separately for those strips only so that the common code stays small.
template <bool TAIL>
__global__ void __launch_bounds__(GD_WARPS * 32, GD_MIN_CTAS) gauss_decide_kernel(FrontParams p, int first_strip, int n_strips, int n_chunks, int rows_per_chunk)
{
    __shared__ __align__(16) float ring_all[GD_WARPS][GD_RING * GD_ROWF];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t task = (int64_t)blockIdx.x * GD_WARPS + warp;
    const int64_t per_frame = (int64_t)n_strips * n_chunks;
    if (task >= per_frame * p.n_frames) return;
    const int f = (int)(task / per_frame);
    const int rr = (int)(task - (int64_t)f * per_frame);
    const int strip = first_strip + rr % n_strips, chunk = rr / n_strips;
    const int y0 = chunk * rows_per_chunk, y1 = min(p.h, y0 + rows_per_chunk);
    float *ring = ring_all[warp];

    const int xs = strip * 128, x0 = xs + 4 * lane;
    const int pitch = p.pitch;
    const float c_mask = GD_MAGIC - (float)p.t_mask, c_mark = GD_MAGIC - (float)p.t_marker;
    const bool row_tail = TAIL && x0 >= p.row_tail_from, col_tail = TAIL && x0 >= p.col_tail_from;
    const uint8_t *src = p.plane + (int64_t)f * p.plane_stride + (int64_t)y0 * pitch + 8 + x0;      // row y0 - 5
    // halo words of a row: lane & 3 -> columns xs-8, xs-4, xs+128, xs+132 ; lane >> 2 -> row within the group of 8
    const int hj = lane & 3, hk = lane >> 2;
    const uint8_t *hsrc = p.plane + (int64_t)f * p.plane_stride + (int64_t)(y0 + hk) * pitch + 8 + xs + (hj < 2 ? 4 * hj - 8 : 120 + 4 * hj);
    float *hdst = ring + hk * GD_ROWF + (hj < 2 ? 4 * hj : 128 + 4 * hj);
    uint8_t *dst = p.decisions + (int64_t)f * p.dec_stride + (int64_t)(y0 - 10) * p.dec_pitch + strip * 32 + lane;
    const int dec_pitch = p.dec_pitch;
    const int n_steps = (y1 - y0) + 10;

    float2 win[11][2];
#pragma unroll
    for (int i = 0; i < 11; ++i) win[i][0] = win[i][1] = make_float2(0.f, 0.f);

    // main words travel four rows (two iterations) ahead of their use
    uint32_t u0 = __ldg(reinterpret_cast<const uint32_t *>(src));
    uint32_t u1 = n_steps > 1 ? __ldg(reinterpret_cast<const uint32_t *>(src + pitch)) : 0u;
    uint32_t u2 = n_steps > 2 ? __ldg(reinterpret_cast<const uint32_t *>(src + 2 * (int64_t)pitch)) : 0u;
    uint32_t u3 = n_steps > 3 ? __ldg(reinterpret_cast<const uint32_t *>(src + 3 * (int64_t)pitch)) : 0u;
    src += 4 * (int64_t)pitch;
    // halo words travel one group of 8 rows ahead of their use, like the main words travel two rows ahead
    uint32_t hw = hk < n_steps ? __ldg(reinterpret_cast<const uint32_t *>(hsrc)) : 0u;
    hsrc += GD_RING * (int64_t)pitch;
    // Two steps per iteration: the step counter, the window-slot dispatch and the pointer updates are paid once per two
    // rows.  (n_steps may be odd: the second half of the last iteration then works on a stale row and stores nothing.)
    int j = 0;
    for (int s = 0; s < n_steps; s += 2) {
        const int slot = s & (GD_RING - 1);       // even, so slot + 1 is inside the ring
        if (slot == 0) {
            __syncwarp();                         // the row pass of the previous step has read the halo zones
            *reinterpret_cast<float4 *>(hdst) = u8x4_to_float4(hw);
            if (s + GD_RING + hk < n_steps) hw = __ldg(reinterpret_cast<const uint32_t *>(hsrc));
            hsrc += GD_RING * (int64_t)pitch;
        }
        const uint32_t ua = u0, ub = u1;
        u0 = u2; u1 = u3;
        if (s + 4 < n_steps) u2 = __ldg(reinterpret_cast<const uint32_t *>(src));
        if (s + 5 < n_steps) u3 = __ldg(reinterpret_cast<const uint32_t *>(src + pitch));
        src += 2 * (int64_t)pitch;
        float *row0 = ring + slot * GD_ROWF, *row1 = row0 + GD_ROWF;
        *reinterpret_cast<float4 *>(row0 + 8 + 4 * lane) = u8x4_to_float4(ua);
        *reinterpret_cast<float4 *>(row1 + 8 + 4 * lane) = u8x4_to_float4(ub);
        __syncwarp();
        float r0[4], r1[4], m0[4], m1[4];
        {
            float a[20];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float4 t4 = *reinterpret_cast<const float4 *>(row0 + 4 * lane + 4 * q);
                a[4 * q] = t4.x; a[4 * q + 1] = t4.y; a[4 * q + 2] = t4.z; a[4 * q + 3] = t4.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) r0[k] = gauss_row<TAIL>(&a[3 + k], row_tail);
        }
        {
            float a[20];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const float4 t4 = *reinterpret_cast<const float4 *>(row1 + 4 * lane + 4 * q);
                a[4 * q] = t4.x; a[4 * q + 1] = t4.y; a[4 * q + 2] = t4.z; a[4 * q + 3] = t4.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) r1[k] = gauss_row<TAIL>(&a[3 + k], row_tail);
        }
        // blurred values of the two output rows (written 5 steps before by this very lane)
        const float4 bf0 = *reinterpret_cast<const float4 *>(ring + ((s + 3) & (GD_RING - 1)) * GD_ROWF + 8 + 4 * lane);
        const float4 bf1 = *reinterpret_cast<const float4 *>(ring + ((s + 4) & (GD_RING - 1)) * GD_ROWF + 8 + 4 * lane);
#define GD_CASE(J) case J: gd_window<J, TAIL>(win, r0, m0, col_tail); gd_window<(J + 1) % 11, TAIL>(win, r1, m1, col_tail); break;
        switch (j) {
            GD_CASE(0) GD_CASE(1) GD_CASE(2) GD_CASE(3) GD_CASE(4) GD_CASE(5) GD_CASE(6) GD_CASE(7) GD_CASE(8) GD_CASE(9)
            default: gd_window<10, TAIL>(win, r0, m0, col_tail); gd_window<0, TAIL>(win, r1, m1, col_tail); break;
        }
#undef GD_CASE
        // keep the row results in their own registers across the merge: otherwise one window slot shares them in one case and
        // every other case pays moves to relocate it (seen in the SASS)
        asm volatile("" : "+f"(r0[0]), "+f"(r0[1]), "+f"(r0[2]), "+f"(r0[3]), "+f"(r1[0]), "+f"(r1[1]), "+f"(r1[2]), "+f"(r1[3]));
        const uint32_t bits0 = gd_decide(m0, bf0, c_mask, c_mark);
        const uint32_t bits1 = gd_decide(m1, bf1, c_mask, c_mark);
        if (s >= 10) *dst = (uint8_t)bits0;
        if (s + 1 >= 10 && s + 1 < n_steps) dst[dec_pitch] = (uint8_t)bits1;
        dst += 2 * dec_pitch;
        j += 2; j = j >= 11 ? j - 11 : j;
    }
}

// mean/std mode (track_eval.py:248-253, cv2.threshold on the blurred image): decision byte = (b > T) per pixel, no Gaussian.
__global__ void __launch_bounds__(256) scalar_decide_kernel(FrontParams p)
{
    const int nb = p.dec_pitch;                   // bytes per decision row (= 4-pixel groups)
    const int64_t total = (int64_t)p.n_frames * p.h * nb;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % nb);
        const int64_t fr = i / nb;
        const int y = (int)(fr % p.h), f = (int)(fr / p.h);
        const int thr = p.scalar_thr[f];
        const uint32_t u = *reinterpret_cast<const uint32_t *>(p.plane + (int64_t)f * p.plane_stride + (int64_t)(y + 5) * p.pitch + 8 + 4 * g);
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) bits |= ((int)((u >> (8 * k)) & 0xFFu) > thr) ? (1u << k) : 0u;
        p.decisions[(int64_t)f * p.dec_stride + (int64_t)y * nb + g] = (uint8_t)bits;
    }
}

// ---- K1c: decision bytes -> bit masks ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t squeeze_nibbles(uint32_t v)   // nibbles at bits 0-3, 8-11, 16-19, 24-27 -> low 16 bits
{
    v = (v | (v >> 4)) & 0x00FF00FFu;
    return (v | (v >> 8)) & 0x0000FFFFu;
}

__global__ void __launch_bounds__(256) pack_masks_kernel(FrontParams p)
{
    // grid = (word pairs of a frame / 256, frames): 32-bit index arithmetic only; one thread packs two adjacent words of a
    // row from 16 decision bytes (dec_pitch is a multiple of 32, so the 16-byte load is aligned and inside the row)
    const int ww = p.ww, f = blockIdx.y;
    const int pairs = (ww + 1) >> 1;
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= (uint32_t)(p.h * pairs)) return;
    const uint32_t y = i / (uint32_t)pairs, word = 2u * (i - y * (uint32_t)pairs);
    const uint32_t inv = p.inverted ? 0xFFFFFFFFu : 0u;
    const uint4 d = *reinterpret_cast<const uint4 *>(p.decisions + (int64_t)f * p.dec_stride + (int64_t)y * p.dec_pitch + 8 * word);
    const uint32_t lo[2] = {d.x, d.z}, hi[2] = {d.y, d.w};
    const int64_t o = ((int64_t)f * p.h + y) * ww + word;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if ((int)word + k >= ww) break;
        const uint32_t mask = squeeze_nibbles(lo[k] & 0x0F0F0F0Fu) | (squeeze_nibbles(hi[k] & 0x0F0F0F0Fu) << 16);
        const uint32_t mark = squeeze_nibbles((lo[k] >> 4) & 0x0F0F0F0Fu) | (squeeze_nibbles((hi[k] >> 4) & 0x0F0F0F0Fu) << 16);
        const int left = p.w - 32 * ((int)word + k);
        const uint32_t valid = left >= 32 ? 0xFFFFFFFFu : ((1u << left) - 1u);
        p.mask_bits[o + k] = (mask ^ inv) & valid;
        if (p.marker_bits) p.marker_bits[o + k] = (mark ^ inv) & valid;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Small helpers used by the debug path and the mean/std mode
// ---------------------------------------------------------------------------------------------------------------------
__global__ void unpack_bits_kernel(const uint32_t *bits, uint8_t *bytes, int64_t rows, int w, int ww)
{
    const int64_t total = rows * w;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / w;
        const int x = (int)(i - r * w);
        bytes[i] = (bits[r * ww + (x >> 5)] >> (x & 31)) & 1u ? 255 : 0;
    }
}

// Per-frame sum and sum of squares of the GREY image (cv2.meanStdDev, track_eval.py:221); exact in u64.
template <int C>
__global__ void __launch_bounds__(256) frame_moments_kernel(const uint8_t *frames, int64_t frame_stride, int h, int w,
                                                            unsigned long long *sums /* [n_frames][2] */)
{
    const int f = blockIdx.y;
    const uint8_t *frame = frames + (int64_t)f * frame_stride;
    const int64_t n = (int64_t)h * w;
    unsigned long long s1 = 0, s2 = 0;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        uint32_t g;
        if (C == 3) g = luma(frame[3 * i], frame[3 * i + 1], frame[3 * i + 2]);
        else g = frame[i];
        s1 += g; s2 += g * g;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[2 * f], s1);
        atomicAdd(&sums[2 * f + 1], s2);
    }
}

// Moving-average threshold of the mean/std mode (track_eval.py:221-238), one thread per frame.
//   hist[-(window)..-1] are the per-frame values of the frames preceding this chunk (oldest first), n_hist of them valid.
//   value_t = mean_t +- std_t +- signed_offset (float64; population std like cv2.meanStdDev)
//   thr_t   = int( sum(values[max(0, t - window) .. t]) / count )      with window = number of frames kept = floor(5*fps)+1
__global__ void moving_threshold_kernel(const unsigned long long *sums, int n_frames, double n_px, int white_on_dark,
                                        double offset, int first_frame, int window, double *values /* [window + n_frames] */,
                                        int32_t *thr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_frames) return;
    // own value (every thread computes its own; values[] of earlier frames of this chunk are recomputed locally to avoid
    // an inter-block dependency)
    // cv2.meanStdDev (probed against 4.13): scale = 1/N; mean = s1*scale; var = s2*scale - mean*mean; std = sqrt(max(var,0))
    const double scale = 1.0 / n_px;
    auto value_of = [&](int u) -> double {
        const double s1 = (double)sums[2 * u], s2 = (double)sums[2 * u + 1];
        const double mean = __dmul_rn(s1, scale);
        double var = __dsub_rn(__dmul_rn(s2, scale), __dmul_rn(mean, mean));
        if (var < 0) var = 0;
        const double sd = sqrt(var);
        // `offset` arrives already sign-flipped for dark-on-light (track_eval.py:132), as the reference has it here
        return white_on_dark ? __dadd_rn(__dadd_rn(mean, sd), offset) : __dsub_rn(__dsub_rn(mean, sd), offset);
    };
    values[window + t] = value_of(t);
    const int g = first_frame + t;                       // global frame index
    int start = g - (window - 1); if (start < 0) start = 0;
    double acc = 0.0; int cnt = 0;
    for (int u = start; u <= g; ++u) {
        const int local = u - first_frame;               // < 0: from history
        const double v = local >= 0 ? value_of(local) : values[window + local];
        acc = cnt == 0 ? v : acc + v;
        ++cnt;
    }
    thr[t] = (int32_t)(acc / (double)cnt);               // int() truncates toward zero
}

__global__ void shift_history_kernel(double *values, int window, int n_frames)
{
    // keep the last `window` values as history for the next chunk: values[i] = values[i + n_frames]
    // (single block, two-phase through registers to be safe for overlapping ranges)
    extern __shared__ double tmp[];
    for (int i = threadIdx.x; i < window; i += blockDim.x) tmp[i] = values[i + n_frames];
    __syncthreads();
    for (int i = threadIdx.x; i < window; i += blockDim.x) values[i] = tmp[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------------------------------
cudaError_t launch_frontend_tile(const FrontParams &p, cudaStream_t st)
{
    dim3 grid((p.w + TILE_W - 1) / TILE_W, (p.h + TILE_H - 1) / TILE_H, p.n_frames);
    if (p.channels == 3) frontend_tile_kernel<3><<<grid, 256, 0, st>>>(p);
    else frontend_tile_kernel<1><<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// K1a + margins (2 launches)
cudaError_t launch_blur_prepass(const FrontParams &p, cudaStream_t st)
{
    const int n_strips = (p.w + PB_COLS - 1) / PB_COLS;
    const int n_chunks = (p.h + 47) / 48;
    const int rows_per_chunk = (p.h + n_chunks - 1) / n_chunks;
    const int64_t tasks = (int64_t)n_strips * n_chunks * p.n_frames;
    const unsigned grid = (unsigned)((tasks + PB_WARPS - 1) / PB_WARPS);
    const bool fast = (p.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.frames) & 3) == 0) && (p.frame_stride % 4 == 0);
    if (p.channels == 3) {
        if (fast) blur_prepass_kernel<3, true><<<grid, PB_WARPS * 32, 0, st>>>(p, n_strips, n_chunks, rows_per_chunk);
        else blur_prepass_kernel<3, false><<<grid, PB_WARPS * 32, 0, st>>>(p, n_strips, n_chunks, rows_per_chunk);
    } else {
        if (fast) blur_prepass_kernel<1, true><<<grid, PB_WARPS * 32, 0, st>>>(p, n_strips, n_chunks, rows_per_chunk);
        else blur_prepass_kernel<1, false><<<grid, PB_WARPS * 32, 0, st>>>(p, n_strips, n_chunks, rows_per_chunk);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    plane_margins_kernel<<<148 * 4, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// K1b: Gaussian + decisions (or the scalar threshold of the mean/std mode).  Strips [0, first_tail) run the tail-free
// kernel on `st`; the strip(s) holding OpenCV's scalar-tail columns run the TAIL variant on `st_tail`, which the caller has
// forked from `st` and joins afterwards, so that the small tail launch fills the gaps of the main launch's last wave.
// *n_launched receives the number of kernels launched.
cudaError_t launch_gauss_decide(const FrontParams &p, cudaStream_t st, cudaStream_t st_tail, int *n_launched)
{
    *n_launched = 0;
    if (p.scalar_thr) {
        scalar_decide_kernel<<<148 * 8, 256, 0, st>>>(p);
        *n_launched = 1;
        return cudaGetLastError();
    }
    const int n_strips = (p.w + 127) / 128;
    const int first_tail = p.col_tail_from < p.w ? p.col_tail_from / 128 : n_strips;     // none if w % 8 == 0
    // row chunks: every chunk recomputes 10 halo rows, every wave of CTAs costs (rows + 10) steps -> pick the chunk count
    // that minimises waves * (rows + 10) for this batch
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, gauss_decide_kernel<false>, GD_WARPS * 32, 0);
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const double slots = (double)sms * ctas_per_sm;
    auto launch = [&](int first_strip, int ns, bool tail, cudaStream_t s_) -> cudaError_t {
        if (ns <= 0) return cudaSuccess;
        ++*n_launched;
        int best_chunks = 1; double best_cost = 1e300;
        for (int nc = 1; nc <= 32 && nc * 16 <= p.h; ++nc) {
            const int rows = (p.h + nc - 1) / nc;
            const double ctas = (double)((int64_t)ns * nc * p.n_frames + GD_WARPS - 1) / GD_WARPS;
            const double waves = ctas / slots;
            const double cost = (waves < 1.0 ? 1.0 : (waves + 0.35)) * (rows + 10);  // +0.35: expected tail of a partial wave
            if (cost < best_cost) { best_cost = cost; best_chunks = nc; }
        }
        const int n_chunks = best_chunks;
        const int rows_per_chunk = (p.h + n_chunks - 1) / n_chunks;
        const int64_t tasks = (int64_t)ns * n_chunks * p.n_frames;
        const unsigned grid = (unsigned)((tasks + GD_WARPS - 1) / GD_WARPS);
        if (tail) gauss_decide_kernel<true><<<grid, GD_WARPS * 32, 0, s_>>>(p, first_strip, ns, n_chunks, rows_per_chunk);
        else gauss_decide_kernel<false><<<grid, GD_WARPS * 32, 0, s_>>>(p, first_strip, ns, n_chunks, rows_per_chunk);
        return cudaGetLastError();
    };
    cudaError_t e = launch(first_tail, n_strips - first_tail, true, st_tail ? st_tail : st);
    if (e != cudaSuccess) return e;
    return launch(0, first_tail, false, st);
}

// K1c (1 launch)
cudaError_t launch_pack_masks(const FrontParams &p, cudaStream_t st)
{
    pack_masks_kernel<<<dim3((unsigned)((p.h * ((p.ww + 1) / 2) + 255) / 256), (unsigned)p.n_frames), 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_unpack_bits(const uint32_t *bits, uint8_t *bytes, int64_t rows, int w, int ww, cudaStream_t st)
{
    unpack_bits_kernel<<<148 * 8, 256, 0, st>>>(bits, bytes, rows, w, ww);
    return cudaGetLastError();
}

cudaError_t launch_frame_moments(const uint8_t *frames, int64_t stride, int n_frames, int h, int w, int channels,
                                 unsigned long long *sums, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * 2 * n_frames, st);
    if (e != cudaSuccess) return e;
    dim3 grid(32, n_frames);
    if (channels == 3) frame_moments_kernel<3><<<grid, 256, 0, st>>>(frames, stride, h, w, sums);
    else frame_moments_kernel<1><<<grid, 256, 0, st>>>(frames, stride, h, w, sums);
    return cudaGetLastError();
}

cudaError_t launch_moving_threshold(const unsigned long long *sums, int n_frames, int h, int w, int white_on_dark,
                                    int offset, int first_frame, int window, double *values, int32_t *thr, cudaStream_t st)
{
    moving_threshold_kernel<<<(n_frames + 127) / 128, 128, 0, st>>>(sums, n_frames, (double)h * (double)w, white_on_dark,
                                                                     (double)offset, first_frame, window, values, thr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    shift_history_kernel<<<1, 256, sizeof(double) * window, st>>>(values, window, n_frames);
    return cudaGetLastError();
}

}  // namespace ysmr
