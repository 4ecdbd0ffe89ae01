// Per-track statistics of evaluate_tracks() on the GPU (partial; see stats.cuh for the columns and the
// reference lines).  One CTA per track: the element-wise columns and both median-filter passes are thread-parallel, the
// largest pairwise distance (scipy pdist + max, O(L^2) per track -- the expensive part of the reference) is spread over the
// CTA and reduced with max, the two Kahan sums pandas defines sequentially are done by one thread each.
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/ysmr_b200.h"
#include "stats.cuh"

namespace ysmr {
namespace {

constexpr int ST_THREADS = 128;
__global__ void __launch_bounds__(ST_THREADS) track_statistics_kernel(const int32_t *track_start, int n_tracks, int64_t n, StatCols c, StatCfg g,
                                                                      uint8_t *ma, uint8_t *mb, double *xn, double *yn, double *out)
{
    __shared__ double red_d[ST_THREADS / 32];
    __shared__ long long red_m[ST_THREADS / 32];
    __shared__ double sh_dist; __shared__ float sh_len;
    const int tr = blockIdx.x, tid = threadIdx.x;
    const int lo = track_start[tr], hi = (tr + 1 < n_tracks ? track_start[tr + 1] : (int)n) - 1;
    for (int i = lo + tid; i <= hi; i += ST_THREADS) {
        ma[i] = (uint8_t)moving_raw(c, g.px, lo, i);
        xn[i] = (c.x[i] - c.x[lo]) / g.px; yn[i] = (c.y[i] - c.y[lo]) / g.px;      // x_norm, y_norm (:963-964)
    }
    __syncthreads();
    for (int i = lo + tid; i <= hi; i += ST_THREADS) mb[i] = (uint8_t)medfilt_bit(ma, lo, hi, i, 3);
    __syncthreads();
    long long motile = 0;
    for (int i = lo + tid; i <= hi; i += ST_THREADS) motile += medfilt_bit(mb, lo, hi, i, g.kernel2);
    // largest squared pairwise distance: row i against the rows behind it, rows dealt round robin
    double max_d2 = 0.0;
    for (int i = lo + tid; i <= hi; i += ST_THREADS) {
        const double xi = xn[i], yi = yn[i];
        for (int j = i + 1; j <= hi; ++j) {
            const double dx = xi - xn[j], dy = yi - yn[j];
            const double d2 = dx * dx + dy * dy;
            max_d2 = d2 > max_d2 ? d2 : max_d2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, max_d2, o);
        max_d2 = od > max_d2 ? od : max_d2;
        motile += __shfl_xor_sync(0xffffffffu, motile, o);
    }
    if ((tid & 31) == 0) { red_d[tid >> 5] = max_d2; red_m[tid >> 5] = motile; }
    if (tid == 32) sh_dist = kahan_distance(c, g.px, lo, hi);            // the two sequential sums, on two different warps
    if (tid == 64) sh_len = kahan_bac_length(c, g.px, lo, hi);
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < ST_THREADS / 32; ++w) { max_d2 = red_d[w] > max_d2 ? red_d[w] : max_d2; motile += red_m[w]; }
        finish_statistics(c, g, lo, hi, sh_dist, max_d2, motile, sh_len, out + (int64_t)tr * STAT_COLUMNS);
    }
}

struct SBufs {
    std::vector<void *> p;
    ~SBufs() { for (void *q : p) cudaFree(q); }
    template <class T> cudaError_t get(T **out, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) { p.push_back(q); *out = (T *)q; }
        return e;
    }
};
thread_local std::string g_stat_error;

}  // namespace
}  // namespace ysmr

using namespace ysmr;

#define TCU(expr)                                                                                    \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            g_stat_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
            return YSMR_E_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

extern "C" {

const char *ysmr_statistics_last_error(void) { return g_stat_error.c_str(); }

int ysmr_track_statistics(int device, int64_t n_rows, const uint32_t *h_track_id, const uint32_t *h_t, const double *h_x, const double *h_y,
                          const double *h_w, const double *h_h, double px_per_um, double fps, int median_kernel,
                          const int32_t *h_track_start, int32_t n_tracks, double *h_stats)
{
    if (!h_track_id || !h_t || !h_x || !h_y || !h_w || !h_h || !h_track_start || !h_stats || n_rows <= 0 || n_rows > 0x7fffffff ||
        n_tracks <= 0 || !(px_per_um > 0.0) || !(fps > 0.0) || median_kernel < 1 || (median_kernel & 1) == 0) {
        g_stat_error = "ysmr_track_statistics: bad argument";
        return YSMR_E_INVALID;
    }
    TCU(cudaSetDevice(device));
    SBufs B;
    uint32_t *t = nullptr; double *x = nullptr, *y = nullptr, *w = nullptr, *h = nullptr, *xn = nullptr, *yn = nullptr, *out = nullptr;
    uint8_t *ma = nullptr, *mb = nullptr; int32_t *starts = nullptr;
    const size_t n = (size_t)n_rows;
    TCU(B.get(&t, n)); TCU(B.get(&x, n)); TCU(B.get(&y, n)); TCU(B.get(&w, n)); TCU(B.get(&h, n)); TCU(B.get(&xn, n)); TCU(B.get(&yn, n));
    TCU(B.get(&ma, n)); TCU(B.get(&mb, n)); TCU(B.get(&starts, (size_t)n_tracks)); TCU(B.get(&out, (size_t)n_tracks * STAT_COLUMNS));
    TCU(cudaMemcpy(t, h_t, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    TCU(cudaMemcpy(x, h_x, n * sizeof(double), cudaMemcpyHostToDevice));
    TCU(cudaMemcpy(y, h_y, n * sizeof(double), cudaMemcpyHostToDevice));
    TCU(cudaMemcpy(w, h_w, n * sizeof(double), cudaMemcpyHostToDevice));
    TCU(cudaMemcpy(h, h_h, n * sizeof(double), cudaMemcpyHostToDevice));
    TCU(cudaMemcpy(starts, h_track_start, (size_t)n_tracks * sizeof(int32_t), cudaMemcpyHostToDevice));
    StatCols c{t, x, y, w, h};
    StatCfg g{px_per_um, fps, median_kernel};
    track_statistics_kernel<<<n_tracks, ST_THREADS>>>(starts, n_tracks, n_rows, c, g, ma, mb, xn, yn, out);
    TCU(cudaGetLastError());
    TCU(cudaMemcpy(h_stats, out, (size_t)n_tracks * STAT_COLUMNS * sizeof(double), cudaMemcpyDeviceToHost));
    (void)h_track_id;
    return YSMR_OK;
}

}  // extern "C"
