// Track selection on the GPU: replaces the data-parallel body of select_tracks(),
// /root/reference/ysmr/track_eval.py:541-843, on the rows of <video>_list.csv (grouped by TRACK_ID, POSITION_T).
//
//   initial clean-up (:609-672)   area = W*H, per-track median (radix select, one CTA per track), first/last frame ->
//                                 per-row keep flag; order-preserving compaction (block counts -> scan -> scatter)
//   statistics (:698-741)         ratio_wh; area quantiles and the motility outer fence from exact order statistics
//                                 (device-wide radix select over order-preserving 64-bit keys, 8 passes of 8 bits);
//                                 per-row distance and outlier flag
//   fine selection (:747-794)     find_good_tracks: one warp per track (select.cuh), recursion as an explicit stack
//   result (:822-835)             per-row flags back on the caller's row numbering
//
// The host (ysmr_b200/select.py) does what is scalar: numpy's linear interpolation between the two order statistics of a
// quantile, the fence, the "too many outliers" switch, logging, the DataFrame.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/ysmr_b200.h"
#include "select.cuh"

namespace ysmr {
namespace {

constexpr int CB = 1024;                 // rows per block of the compaction kernels

__device__ __forceinline__ unsigned long long f64_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k)
{
    const unsigned long long b = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
    return __longlong_as_double((long long)b);
}

__global__ void start_flags_kernel(const uint32_t *track, int64_t n, uint8_t *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || track[i] != track[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(CB) count_flags_kernel(const uint8_t *flag, int64_t n, uint32_t *block_count)
{
    const int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
    const int c = __syncthreads_count(i < n && flag[i]);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
}
// exclusive scan of the block counts in place, by one block; *total = number of flagged rows
__global__ void __launch_bounds__(1024) scan_counts_kernel(uint32_t *block_count, int nb, int64_t *total)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nb ? block_count[i] : 0u;
        uint32_t incl = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = ws[lane];
            uint32_t wi = w;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
            ws[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t excl = carry + ws[warp] + incl - v;
        if (i < nb) block_count[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = (int64_t)carry_s;
}
__global__ void __launch_bounds__(CB) scatter_flags_kernel(const uint8_t *flag, int64_t n, const uint32_t *block_off, int32_t *out_idx)
{
    __shared__ uint32_t wcnt[32];
    const int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool f = i < n && flag[i];
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wcnt[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = wcnt[lane];
        uint32_t wi = w;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
        wcnt[lane] = wi - w;
    }
    __syncthreads();
    if (f) out_idx[block_off[blockIdx.x] + wcnt[warp] + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
}

// k-th smallest (0-based) of the keys a CTA produces with `key_of(i)`, i in [lo, hi): MSB-first radix select, 8 bits a pass
template <class KeyOf>
__device__ unsigned long long cta_radix_select(KeyOf key_of, int lo, int hi, int k, uint32_t *hist, unsigned long long *sh_state)
{
    unsigned long long prefix = 0ull, mask = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0u;
        __syncthreads();
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const unsigned long long key = key_of(i);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255ull], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int cum = 0, b = 0;
            for (; b < 255; ++b) { if (cum + (int)hist[b] > k) break; cum += (int)hist[b]; }
            sh_state[0] = (unsigned long long)b; sh_state[1] = (unsigned long long)(k - cum);
        }
        __syncthreads();
        prefix |= sh_state[0] << shift; mask |= 255ull << shift; k = (int)sh_state[1];
        __syncthreads();
    }
    return prefix;
}

// Initial clean-up of one track (track_eval.py:609-660): median of the area over ALL rows of the track (pandas
// group_median_float64: the middle value, or the mean of the two middle values), length = last - first frame + 1 as
// uint16, then the per-row keep flag and the area column.
struct CleanParams { double area_lo, area_hi, area_factor; int min_len; };
__global__ void __launch_bounds__(128) track_clean_kernel(const int32_t *track_start, int n_tracks, int64_t n, const uint32_t *t,
                                                          const double *w, const double *h, CleanParams p, double *area, uint8_t *keep)
{
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long st[2];
    const int tr = blockIdx.x;
    const int lo = track_start[tr], hi = tr + 1 < n_tracks ? track_start[tr + 1] : (int)n;
    const int cnt = hi - lo;
    auto key_of = [&](int i) { return f64_key(w[i] * h[i]); };
    double med = key_f64(cta_radix_select(key_of, lo, hi, cnt / 2, hist, st));
    if ((cnt & 1) == 0) {
        const double below = key_f64(cta_radix_select(key_of, lo, hi, cnt / 2 - 1, hist, st));
        med = (med + below) / 2.0;
    }
    const uint32_t len16 = (t[hi - 1] - t[lo] + 1u) & 0xFFFFu;                   // .astype(np.uint16), track_eval.py:645-647
    const bool track_ok = med >= p.area_lo && med <= p.area_hi && (int)len16 >= p.min_len;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double a = w[i] * h[i];
        bool ok = track_ok && a != 0.0;
        if (p.area_factor != 0.0) ok = ok && a <= med * p.area_factor;
        area[i] = a;
        keep[i] = ok ? 1 : 0;
    }
}

__global__ void gather_kernel(const int32_t *idx, int64_t n2, const uint32_t *track, const uint32_t *t, const double *x, const double *y,
                              const double *w, const double *h, const double *area, uint32_t *track2, uint32_t *t2, double *x2, double *y2,
                              double *area2, double *ratio2)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n2) return;
    const int i = idx[j];
    track2[j] = track[i]; t2[j] = t[i]; x2[j] = x[i]; y2[j] = y[i]; area2[j] = area[i];
    const double wv = w[i], hv = h[i];
    ratio2[j] = hv <= wv ? hv / wv : wv / hv;                                    // track_eval.py:698
}

// df['distance'] before the fence test (track_eval.py:716-718): 0 at track starts
__global__ void distance_kernel(const uint8_t *start_flag, const uint32_t *t, const double *x, const double *y, int64_t n, double *dist,
                                unsigned long long *keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0.0;
    if (!start_flag[i]) {
        const double dx = x[i] - x[i - 1], dy = y[i] - y[i - 1];
        const double dt = (double)t[i] - (double)t[i - 1];
        d = sqrt(dx * dx + dy * dy) / dt;
    }
    dist[i] = d;
    keys[i] = f64_key(d);
}
__global__ void keys_kernel(const double *v, int64_t n, unsigned long long *keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = f64_key(v[i]);
}

// Device-wide radix select.  state = {prefix, mask, k}; one histogram pass + one pick per 8 bits.
__global__ void __launch_bounds__(256) select_hist_kernel(const unsigned long long *keys, int64_t n, const unsigned long long *state, int shift,
                                                          uint32_t *hist)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long prefix = state[0], mask = state[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if ((key & mask) == prefix) atomicAdd(&sh[(key >> shift) & 255ull], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void select_pick_kernel(uint32_t *hist, unsigned long long *state, int shift, double *out)
{
    if (threadIdx.x == 0) {
        unsigned long long k = state[2], cum = 0ull;
        int b = 0;
        for (; b < 255; ++b) { if (cum + hist[b] > k) break; cum += hist[b]; }
        state[0] |= (unsigned long long)b << shift; state[1] |= 255ull << shift; state[2] = k - cum;
        if (shift == 0) *out = key_f64(state[0]);
    }
    __syncthreads();
    hist[threadIdx.x] = 0u;
}

__global__ void outlier_kernel(const double *dist, int64_t n, double fence, int8_t *outl, unsigned long long *count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool o = i < n && dist[i] > fence;
    if (i < n) outl[i] = o ? 1 : 0;
    const unsigned m = __ballot_sync(0xffffffffu, o);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

struct WarpLanes {
    __device__ int lane() const { return threadIdx.x & 31; }
    __device__ int lanes() const { return 32; }
    __device__ void max_first(double &v, int &i) const
    {
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
    }
    __device__ int sum(int v) const { return __reduce_add_sync(0xffffffffu, v); }
    __device__ int min_i(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    __device__ double min_d(double v) const
    {
        for (int o = 16; o > 0; o >>= 1) { const double ov = __shfl_xor_sync(0xffffffffu, v, o); v = ov < v ? ov : v; }
        return v;
    }
    __device__ double max_d(double v) const
    {
        for (int o = 16; o > 0; o >>= 1) { const double ov = __shfl_xor_sync(0xffffffffu, v, o); v = ov > v ? ov : v; }
        return v;
    }
    __device__ double bcast(double v) const { return __shfl_sync(0xffffffffu, v, 0); }
};

extern __shared__ __align__(16) unsigned char ysmr_select_smem[];
__global__ void __launch_bounds__(32) select_tracks_kernel(const int32_t *track_start, int n_tracks, int64_t n, SelectCols c, SelectCfg g,
                                                           int stack_cap, const int32_t *orig_idx, uint8_t *good, unsigned long long *kick_hist,
                                                           unsigned long long *n_good_tracks)
{
    SelectSeg *stack = reinterpret_cast<SelectSeg *>(ysmr_select_smem);
    const int tr = blockIdx.x;
    const int lo = track_start[tr], hi = (tr + 1 < n_tracks ? track_start[tr + 1] : (int)n) - 1;
    int gs = -1, ge = -1;
    const WarpLanes wp;
    const int kick = select_track(wp, c, g, lo, hi, stack, stack_cap, &gs, &ge);
    if (threadIdx.x == 0) {
        atomicAdd(&kick_hist[kick], 1ull);
        if (gs >= 0) atomicAdd(n_good_tracks, 1ull);
    }
    if (gs >= 0)
        for (int i = gs + threadIdx.x; i <= ge; i += 32) good[orig_idx[i]] = 1;
}

__global__ void scatter_clean_kernel(const int32_t *orig_idx, int64_t n2, int32_t *clean_index)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n2) clean_index[orig_idx[j]] = (int32_t)j;
}

struct Bufs {
    std::vector<void *> p;
    ~Bufs() { for (void *q : p) cudaFree(q); }
    template <class T> cudaError_t get(T **out, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) { p.push_back(q); *out = (T *)q; }
        return e;
    }
};

thread_local std::string g_select_error;

}  // namespace
}  // namespace ysmr

using namespace ysmr;

#define SCU(expr)                                                                                      \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            g_select_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
            return YSMR_E_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

namespace {
int g_select_launches = 0;

// order-preserving compaction of the flagged rows: out_idx[0 .. *count) (device), count copied to the host
int compact(const uint8_t *flag, int64_t n, uint32_t *block_count, int64_t *d_total, int32_t *out_idx, int64_t *h_total, cudaStream_t st)
{
    const int nb = (int)((n + CB - 1) / CB);
    count_flags_kernel<<<nb, CB, 0, st>>>(flag, n, block_count);
    scan_counts_kernel<<<1, 1024, 0, st>>>(block_count, nb, d_total);
    scatter_flags_kernel<<<nb, CB, 0, st>>>(flag, n, block_count, out_idx);
    g_select_launches += 3;
    SCU(cudaGetLastError());
    SCU(cudaMemcpyAsync(h_total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SCU(cudaStreamSynchronize(st));
    return YSMR_OK;
}

// the k-th smallest (0-based) of n keys -> *d_out
int radix_select(const unsigned long long *keys, int64_t n, int64_t k, unsigned long long *state, uint32_t *hist, double *d_out, cudaStream_t st)
{
    const unsigned long long init[3] = {0ull, 0ull, (unsigned long long)k};
    SCU(cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, st));
    SCU(cudaMemsetAsync(hist, 0, 256 * sizeof(uint32_t), st));
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    for (int shift = 56; shift >= 0; shift -= 8) {
        select_hist_kernel<<<grid, 256, 0, st>>>(keys, n, state, shift, hist);
        select_pick_kernel<<<1, 256, 0, st>>>(hist, state, shift, d_out);
        g_select_launches += 2;
    }
    SCU(cudaGetLastError());
    return YSMR_OK;
}
}  // namespace

extern "C" {

const char *ysmr_select_last_error(void) { return g_select_error.c_str(); }

int ysmr_select_tracks(int device, int64_t n_rows, const uint32_t *h_track_id, const uint32_t *h_t, const double *h_x, const double *h_y,
                       const double *h_w, const double *h_h, const ysmr_select_params *sp, uint8_t *h_good, int32_t *h_clean_index,
                       int64_t *h_kick_reasons, double *h_info)
{
    if (!sp || !h_track_id || !h_t || !h_x || !h_y || !h_w || !h_h || !h_good || !h_clean_index || !h_kick_reasons || !h_info ||
        n_rows < 0 || n_rows > 0x7fffffff) {
        g_select_error = "ysmr_select_tracks: bad argument";
        return YSMR_E_INVALID;
    }
    if (sp->max_recursion < 0 || sp->max_recursion > 4000) {
        g_select_error = "ysmr_select_tracks: 'maximal recursion depth' outside 0 .. 4000";
        return YSMR_E_INVALID;
    }
    g_select_launches = 0;
    for (int i = 0; i < 9; ++i) h_kick_reasons[i] = 0;
    for (int i = 0; i < YSMR_SELECT_INFO; ++i) h_info[i] = 0.0;
    memset(h_good, 0, (size_t)n_rows);
    for (int64_t i = 0; i < n_rows; ++i) h_clean_index[i] = -1;
    h_info[YSMR_SI_STATUS] = YSMR_SEL_OK;
    h_info[YSMR_SI_ROWS_BEFORE] = (double)n_rows;
    if (n_rows < sp->min_len_frames || n_rows == 0) {           // track_eval.py:599-606
        h_info[YSMR_SI_STATUS] = YSMR_SEL_TOO_SHORT_BEFORE;
        return YSMR_OK;
    }
    SCU(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    const int64_t n = n_rows;
    Bufs B;
    uint32_t *track = nullptr, *t = nullptr, *track2 = nullptr, *t2 = nullptr, *block_count = nullptr, *hist = nullptr;
    double *x = nullptr, *y = nullptr, *w = nullptr, *h = nullptr, *area = nullptr, *x2 = nullptr, *y2 = nullptr, *area2 = nullptr, *ratio2 = nullptr,
           *dist = nullptr, *d_q = nullptr;
    uint8_t *flag = nullptr, *keep = nullptr, *good = nullptr;
    int8_t *outl = nullptr;
    int32_t *starts = nullptr, *cidx = nullptr, *starts2 = nullptr, *clean_index = nullptr;
    int64_t *d_total = nullptr;
    unsigned long long *keys = nullptr, *state = nullptr, *counters = nullptr;
    SCU(B.get(&track, n)); SCU(B.get(&t, n)); SCU(B.get(&x, n)); SCU(B.get(&y, n)); SCU(B.get(&w, n)); SCU(B.get(&h, n));
    SCU(B.get(&area, n)); SCU(B.get(&flag, n)); SCU(B.get(&keep, n)); SCU(B.get(&good, n)); SCU(B.get(&starts, n)); SCU(B.get(&cidx, n));
    SCU(B.get(&clean_index, n));
    SCU(B.get(&block_count, (n + CB - 1) / CB + 1)); SCU(B.get(&d_total, 1)); SCU(B.get(&hist, 256)); SCU(B.get(&state, 3));
    SCU(B.get(&d_q, 8)); SCU(B.get(&counters, 16));
    SCU(cudaMemcpyAsync(track, h_track_id, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SCU(cudaMemcpyAsync(t, h_t, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SCU(cudaMemcpyAsync(x, h_x, n * sizeof(double), cudaMemcpyHostToDevice, st));
    SCU(cudaMemcpyAsync(y, h_y, n * sizeof(double), cudaMemcpyHostToDevice, st));
    SCU(cudaMemcpyAsync(w, h_w, n * sizeof(double), cudaMemcpyHostToDevice, st));
    SCU(cudaMemcpyAsync(h, h_h, n * sizeof(double), cudaMemcpyHostToDevice, st));
    SCU(cudaMemsetAsync(good, 0, n, st));
    SCU(cudaMemsetAsync(counters, 0, 16 * sizeof(unsigned long long), st));
    SCU(cudaMemsetAsync(clean_index, 0xff, n * sizeof(int32_t), st));

    // ---- tracks of the input, initial clean-up
    const int grid_n = (int)((n + 255) / 256);
    start_flags_kernel<<<grid_n, 256, 0, st>>>(track, n, flag);
    ++g_select_launches;
    int64_t n_tracks = 0;
    int rc = compact(flag, n, block_count, d_total, starts, &n_tracks, st);
    if (rc != YSMR_OK) return rc;
    h_info[YSMR_SI_TRACKS_BEFORE] = (double)n_tracks;
    CleanParams cp{sp->area_lo, sp->area_hi, sp->area_factor, sp->min_len_frames};
    track_clean_kernel<<<(int)n_tracks, 128, 0, st>>>(starts, (int)n_tracks, n, t, w, h, cp, area, keep);
    ++g_select_launches;
    int64_t n2 = 0;
    rc = compact(keep, n, block_count, d_total, cidx, &n2, st);
    if (rc != YSMR_OK) return rc;
    h_info[YSMR_SI_ROWS_AFTER] = (double)n2;
    if (n2 < sp->min_len_frames || n2 == 0) {                   // track_eval.py:676-684
        h_info[YSMR_SI_STATUS] = YSMR_SEL_TOO_SHORT_AFTER;
        h_info[YSMR_SI_LAUNCHES] = g_select_launches;
        return YSMR_OK;
    }
    SCU(B.get(&track2, n2)); SCU(B.get(&t2, n2)); SCU(B.get(&x2, n2)); SCU(B.get(&y2, n2)); SCU(B.get(&area2, n2)); SCU(B.get(&ratio2, n2));
    SCU(B.get(&dist, n2)); SCU(B.get(&outl, n2)); SCU(B.get(&keys, n2)); SCU(B.get(&starts2, n2));
    const int grid_2 = (int)((n2 + 255) / 256);
    gather_kernel<<<grid_2, 256, 0, st>>>(cidx, n2, track, t, x, y, w, h, area, track2, t2, x2, y2, area2, ratio2);
    scatter_clean_kernel<<<grid_2, 256, 0, st>>>(cidx, n2, clean_index);
    start_flags_kernel<<<grid_2, 256, 0, st>>>(track2, n2, flag);
    g_select_launches += 3;
    int64_t n_tracks2 = 0;
    rc = compact(flag, n2, block_count, d_total, starts2, &n_tracks2, st);
    if (rc != YSMR_OK) return rc;
    h_info[YSMR_SI_TRACKS_AFTER] = (double)n_tracks2;

    // ---- quantiles: the host interpolates between two exact order statistics (numpy 'linear')
    auto ranks_of = [&](double q, int64_t cnt, int64_t *prev, int64_t *next, double *gamma) {
        // numpy 2.3 _QuantileMethods['linear']: virtual index (n - 1) * q; _get_indexes, _get_gamma
        const double vi = (double)(cnt - 1) * q;
        double pf = floor(vi);
        int64_t p_ = (int64_t)pf, n_ = p_ + 1;
        if (vi >= (double)(cnt - 1)) { p_ = cnt - 1; n_ = cnt - 1; }
        if (vi < 0.0) { p_ = 0; n_ = 0; }
        *prev = p_; *next = n_; *gamma = vi - pf;
    };
    auto lerp = [](double a, double b, double tt) {
        const double d = b - a;
        return tt >= 0.5 ? b - d * (1.0 - tt) : a + d * tt;
    };
    auto quantile_pair = [&](const unsigned long long *kk, int64_t cnt, double qa, double qb, double *ra, double *rb) -> int {
        int64_t pr[2], nx[2]; double gm[2];
        ranks_of(qa, cnt, &pr[0], &nx[0], &gm[0]);
        ranks_of(qb, cnt, &pr[1], &nx[1], &gm[1]);
        for (int i = 0; i < 2; ++i) {
            int r_ = radix_select(kk, cnt, pr[i], state, hist, d_q + 2 * i, st);
            if (r_ != YSMR_OK) return r_;
            r_ = radix_select(kk, cnt, nx[i], state, hist, d_q + 2 * i + 1, st);
            if (r_ != YSMR_OK) return r_;
        }
        double v[4];
        SCU(cudaMemcpyAsync(v, d_q, sizeof(v), cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        *ra = lerp(v[0], v[1], gm[0]); *rb = lerp(v[2], v[3], gm[1]);
        return YSMR_OK;
    };
    double lower = -1.0, upper = INFINITY;
    if (sp->q_area > 0.0) {
        keys_kernel<<<grid_2, 256, 0, st>>>(area2, n2, keys);
        ++g_select_launches;
        rc = quantile_pair(keys, n2, sp->q_area, 1.0 - sp->q_area, &lower, &upper);
        if (rc != YSMR_OK) return rc;
    }
    h_info[YSMR_SI_Q1_AREA] = lower; h_info[YSMR_SI_Q3_AREA] = upper;
    SCU(cudaMemsetAsync(outl, 0, n2, st));
    if (sp->omit_motility_outliers) {
        distance_kernel<<<grid_2, 256, 0, st>>>(flag, t2, x2, y2, n2, dist, keys);
        ++g_select_launches;
        double q1 = 0.0, q3 = 0.0;
        rc = quantile_pair(keys, n2, 0.25, 0.75, &q1, &q3);
        if (rc != YSMR_OK) return rc;
        const double fence = (q3 - q1) * 3.0 + q3;                            // track_eval.py:720
        outlier_kernel<<<grid_2, 256, 0, st>>>(dist, n2, fence, outl, counters + 10);
        ++g_select_launches;
        unsigned long long n_out = 0;
        SCU(cudaMemcpyAsync(&n_out, counters + 10, sizeof(n_out), cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
        h_info[YSMR_SI_Q1_DIST] = q1; h_info[YSMR_SI_Q3_DIST] = q3; h_info[YSMR_SI_FENCE] = fence; h_info[YSMR_SI_OUTLIERS] = (double)n_out;
        if ((double)n_out / (double)n2 > sp->stop_outliers_above) {           // track_eval.py:728-741
            SCU(cudaMemsetAsync(outl, 0, n2, st));
            h_info[YSMR_SI_OUTLIERS_OFF] = 1.0;
        }
    }

    // ---- fine selection: one warp per track
    SelectCols c{t2, x2, y2, area2, ratio2, outl};
    SelectCfg g{};
    g.min_len = sp->min_len_frames; g.max_holes = sp->max_holes; g.max_recursion = sp->max_recursion; g.max_empty = sp->max_empty;
    g.lower = lower; g.upper = upper; g.ratio_min = sp->ratio_min; g.ratio_max = sp->ratio_max; g.edge = sp->edge;
    g.frame_h = sp->frame_h; g.frame_w = sp->frame_w; g.limit_frames = sp->limit_frames; g.limit_exactly = sp->limit_exactly;
    const int stack_cap = sp->max_recursion + 2;
    const size_t smem = (size_t)stack_cap * sizeof(SelectSeg);
    if (smem > 48 * 1024) SCU(cudaFuncSetAttribute(select_tracks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_tracks_kernel<<<(int)n_tracks2, 32, smem, st>>>(starts2, (int)n_tracks2, n2, c, g, stack_cap, cidx, good, counters, counters + 9);
    ++g_select_launches;
    SCU(cudaGetLastError());
    unsigned long long hc[10];
    SCU(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    SCU(cudaMemcpyAsync(h_good, good, n, cudaMemcpyDeviceToHost, st));
    SCU(cudaMemcpyAsync(h_clean_index, clean_index, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SCU(cudaStreamSynchronize(st));
    for (int i = 0; i < 9; ++i) h_kick_reasons[i] = (int64_t)hc[i];
    h_info[YSMR_SI_GOOD_TRACKS] = (double)hc[9];
    h_info[YSMR_SI_LAUNCHES] = g_select_launches;
    if (hc[9] == 0) h_info[YSMR_SI_STATUS] = YSMR_SEL_NO_TRACKS;              // track_eval.py:816-819
    return YSMR_OK;
}

}  // extern "C"
