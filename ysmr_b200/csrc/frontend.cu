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
//   frontend_strip_kernel  -- the production kernel: each warp marches a 128-column strip down the frame with the
//                             11-row column window in registers, 128-bit shared-memory traffic and packed stores.
#include "frontend.cuh"

namespace ysmr {

__device__ __forceinline__ int reflect101(int p, int n)
{
    // one reflection is exact for -n < p < 2n-1 (all positions whose value is ever used: halos of at most 6 with
    // n >= 16); positions further out only occur in tiles hanging over the image edge and are clamped to stay in bounds
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p < 0 ? 0 : (p >= n ? n - 1 : p);
}

__device__ __forceinline__ int clampi(int p, int n) { return p < 0 ? 0 : (p >= n ? n - 1 : p); }

__device__ __forceinline__ uint32_t luma(uint32_t b, uint32_t g, uint32_t r)
{
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

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
