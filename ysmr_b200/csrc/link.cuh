// K4/K5 -- the sequential linker: nearest-neighbour association with disappearance counters and the in-loop
// Gaussian-sum FIR filter.
//
// Replaces CentroidTracker.update (/root/reference/ysmr/tracker.py:93-230) -- including scipy's cdist (tracker.py:151),
// the row-min argsort greedy match (158-189), the two asymmetric branches (198-217) and CPython's set iteration order
// for several births in one frame (193, 216-217) -- and GaussianSumFIR.correct/predict (/root/reference/ysmr/
// gsff.py:204-347), plus the row append of the loop (track_eval.py:313-316).  Data-parallel restatement:
// `rows` of the reference is a permutation, so "row in used_rows" never fires and the greedy loop reduces to: every
// detection is won by the track with the smallest row-minimum among the tracks whose nearest detection it is.
//
// One CTA walks the frames of a chunk in order; within a frame the phases below are loops over tracks/detections
// separated by cta.sync().  The Cta policy supplies barrier, scan and atomics (device: link.cu; host emulation for
// the CPU test-suite: tests/host_emul).  All association and filter arithmetic is float64 like the reference.
#pragma once
#include "common.cuh"

#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

namespace ysmr {

constexpr int LINK_MAX_FILTERS = 4;
constexpr int LINK_MAX_HORIZON = 64;

struct LinkConfig {
    double max_disappeared;              // = fps (tracker.py:42, track_eval.py:110)
    double max_distance;                 // <= 0: no gate (reference behaviour)
    int use_gsff, n_f;
    int n_i[LINK_MAX_FILTERS];           // horizons (gsff.py:103-109)
    int hist_len;                        // n_i[n_f-1] + 1 (gsff.py:317-318)
    int cross_zero;                      // 1: the x<-y / y<-x gain entries are all exactly zero
    int xy_same;                         // 1: the x and y rows of every gain carry identical taps
    // gains, device memory: for filter i four arrays of n_i[i] doubles: xx, xy, yx, yy
    //   x_hat = sum_k xx[k]*mx[k] + xy[k]*my[k] ; y_hat = sum_k yx[k]*mx[k] + yy[k]*my[k]
    const double *gain[LINK_MAX_FILTERS];
    int max_tracks, max_blobs;
    const double *exp_tab;               // np_exp_table_bits as 32 doubles (device memory)
    double fast_gain[3][30];             // lane path: the x-row taps of filters 0..2 (horizons 10 / 20 / 30), as kernel
                                         // PARAMETERS: a multiply-add takes a tap straight from the constant bank
    double grid_cell;                    // general path: cell size of the detection grid (pixels) ...
    int grid_w, grid_h;                  // ... and its extent, grid_w * grid_h <= LINK_GRID_CELLS
};

// Persistent linker state (device memory).  Tracks live in physical slots; `order` lists the slots in insertion order
// (the reference's OrderedDict order), so deregistration only compacts `order`.
struct LinkState {
    int32_t *hdr;                        // [8]: n_tracks, next_id, n_free, order_sel, frames_done, last_n_live
    int32_t *order[2];                   // [max_tracks] ping-pong
    int32_t *free_slots;                 // [max_tracks] stack of free physical slots
    int32_t *id;                         // per slot ...
    double *px, *py;                     //   position used for the next association (GSFF prediction)
    float *iw, *ih, *ideg;               //   additional_info (w, h, deg) or zeros
    int32_t *gone;                       //   consecutive misses
    int32_t *mode, *hist_n;              //   GSFF: active filters, valid history entries
    double *hist;                        //   [hist_len][max_tracks][2], row = frame counter mod hist_len
    double *wgt;                         //   [slot][LINK_MAX_FILTERS]
    double *xh;                          //   [slot][LINK_MAX_FILTERS][2]
};

// Per-frame scratch (device: global memory owned by the context).
struct LinkScratch {
    double *row_min;                     // [max_tracks]
    int32_t *row_arg;                    // [max_tracks]
    uint32_t *flag;                      // [max(max_tracks, max_blobs) + 1] scan buffer
    int32_t *list;                       // [max_blobs] unused detections ascending, then in registration order
    int32_t *table;                      // [set_table_size] CPython set emulation
    int set_table_size;
    long long *phase_cycles;             // [16] (ABI 2 per-phase cycle counters; not maintained by the lane linker)
    // candidate tables of the fast path, rebuilt per launch (link.cu: link_prep_kernel)
    int32_t *succ;                       // [prep_frames][256] detection q of frame t -> nearest detection of frame t+1, or -1
    float *thr2;                         // [prep_frames][256] squared acceptance radius of detection q of frame t
    int prep_frames;                     // frames per sequential launch
    int32_t *lane_done;                  // [1] frames of the launch the fast path handled (read by link_general_kernel)
    float prep_margin;                   // float32 rounding bound of coordinates / distances (pixels)
    int32_t *grid_ws;                    // [8] workspace of the cooperative-grid general path (link.cu: GridLinkCta), zeroed
};

struct RowOut {                          // must match ysmr_row (include/ysmr_b200.h)
    int32_t frame, track_id;
    double x, y;
    float w, h, deg;
    int32_t pad;
};

enum { LINK_ST_TRACK_OVERFLOW = 8, LINK_ST_ROW_OVERFLOW = 16, LINK_ST_GATE_TIMEOUT = 32 };

YSMR_HD unsigned long long f64_bits(double v)
{
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long u; memcpy(&u, &v, 8); return u;
#endif
}

// ---- CPython 3.12 set iteration order of {keys inserted ascending}; SURVEY A.10 / oracle/setorder.py ---------------
// keys[0..n) ascending on entry, registration order on exit.  table: scratch of >= set_table_capacity(n) ints.
YSMR_HD int set_table_capacity(int n)
{
    int size = 8;
    while (size <= 4 * n) size <<= 1;   // the largest table ever built for n keys ...
    return 2 * size;                    // ... plus room for the rebuild (old table in front of the new one)
}

YSMR_HD void set_insert(int32_t *table, int mask, int key)
{
    unsigned perturb = (unsigned)key;
    int i = key & mask;
    for (;;) {
        const int last = (i + 9 <= mask) ? i + 9 : i;
        for (int j = i; j <= last; ++j)
            if (table[j] < 0) { table[j] = key; return; }
        perturb >>= 5;
        i = (int)(((unsigned)i * 5u + 1u + perturb) & (unsigned)mask);
    }
}

YSMR_HD void cpython_set_order(int32_t *keys, int n, int32_t *table)
{
    if (n <= 1) return;
    int mask = 7, fill = 0;
    for (int j = 0; j <= mask; ++j) table[j] = -1;
    for (int q = 0; q < n; ++q) {
        set_insert(table, mask, keys[q]);
        ++fill;
        if (fill * 5 >= mask * 3) {
            const int want = fill > 50000 ? fill * 2 : fill * 4;
            int size = 8;
            while (size <= want) size <<= 1;
            // rebuild: the old entries are re-inserted in old slot order.  Old table occupies [0, mask]; build the new
            // one behind it, then move it to the front.
            int32_t *nt = table + (mask + 1);
            for (int j = 0; j < size; ++j) nt[j] = -1;
            for (int j = 0; j <= mask; ++j)
                if (table[j] >= 0) set_insert(nt, size - 1, table[j]);
            for (int j = 0; j < size; ++j) table[j] = nt[j];
            mask = size - 1;
        }
    }
    int k = 0;
    for (int j = 0; j <= mask; ++j)
        if (table[j] >= 0) keys[k++] = table[j];
}

// ---- float64 arithmetic with explicit rounding ----------------------------------------------------------------------
// The GSFF recursion feeds its own output back while a track is unmatched, which amplifies a last-bit difference by
// about 2.5x per frame: after a 30-frame coast anything but the reference's exact operation sequence is pixels away and
// a different track wins the next contested detection.  So the filter below is not "the same formula", it is the same
// ROUNDINGS as the reference's NumPy calls on x86-64 (verified bit for bit on the build host, tests/test_np_arith.py):
//   * numpy.dot(gain (4 x 2n), y (2n))  -> OpenBLAS dgemv_t, Haswell kernel (used for every AVX2-or-later x86 target):
//     four accumulators over the positions p = 0..2n-1 (p mod 4), each a chain of FMAs starting from +0, combined as
//     (l0 + l2) + (l1 + l3); if 2n % 4 == 2 the last two positions are added as  + fma(a0, y0, a1 * y1).
//   * numpy.dot(d, d) (2 elements)      -> fma(d1, d1, d0 * d0)
//   * numpy.exp (float64, AVX-512)      -> SVML __svml_exp8_ha: restated in np_exp_nonpos below
//   * sum(list * array), numpy.sum(axis=1) over <= 4 entries -> left to right, products rounded separately
//   * a * b / c                         -> IEEE multiply, then IEEE divide
#if defined(__CUDA_ARCH__)
YSMR_D double d_mul(double a, double b) { return __dmul_rn(a, b); }
YSMR_D double d_add(double a, double b) { return __dadd_rn(a, b); }
YSMR_D double d_sub(double a, double b) { return __dsub_rn(a, b); }
YSMR_D double d_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
YSMR_D double d_fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
YSMR_D double d_div(double a, double b) { return __ddiv_rn(a, b); }
YSMR_D double d_from_bits(unsigned long long u) { return __longlong_as_double((long long)u); }
#else
}  // namespace ysmr
#include <fenv.h>
#include <string.h>
namespace ysmr {
inline double d_mul(double a, double b) { return a * b; }
inline double d_add(double a, double b) { return a + b; }
inline double d_sub(double a, double b) { return a - b; }
inline double d_fma(double a, double b, double c) { return fma(a, b, c); }
inline double d_fma_rz(double a, double b, double c)
{
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    volatile double va = a, vb = b, vc = c;
    volatile double r = fma(va, vb, vc);
    fesetround(old);
    return r;
}
inline double d_div(double a, double b) { return a / b; }
inline double d_from_bits(unsigned long long u) { double v; memcpy(&v, &u, 8); return v; }
#endif

// Table of the exp restatement: 2^(j/16) as high part and RELATIVE low part, j = 0..15 (32 doubles: hi[16], lo[16]).
// Computed to the published construction (hi = 2^(j/16) rounded to nearest, lo = (2^(j/16) - hi) / hi) and checked
// against numpy.exp over 10^6 arguments by tests/test_np_arith.py.
constexpr int NP_EXP_TABLE = 32;
YSMR_HD unsigned long long np_exp_table_bits(int k)
{
    constexpr unsigned long long t[NP_EXP_TABLE] = {
        0x3ff0000000000000ull, 0x3ff0b5586cf9890full, 0x3ff172b83c7d517bull, 0x3ff2387a6e756238ull,
        0x3ff306fe0a31b715ull, 0x3ff3dea64c123422ull, 0x3ff4bfdad5362a27ull, 0x3ff5ab07dd485429ull,
        0x3ff6a09e667f3bcdull, 0x3ff7a11473eb0187ull, 0x3ff8ace5422aa0dbull, 0x3ff9c49182a3f090ull,
        0x3ffae89f995ad3adull, 0x3ffc199bdd85529cull, 0x3ffd5818dcfba487ull, 0x3ffea4afa2a490daull,
        0x0000000000000000ull, 0x3c979aa65d837b6dull, 0xbc801b15eaa59348ull, 0x3c968efde3a8a894ull,
        0x3c834d754db0abb6ull, 0x3c859f48a72a4c6dull, 0x3c7690cebb7aafb0ull, 0x3c9063e1e21c5409ull,
        0xbc93b3efbf5e2228ull, 0xbc7b32dcb94da51dull, 0x3c8db72fc1f0eab4ull, 0x3c71affc2b91ce27ull,
        0x3c8c1a7792cb3387ull, 0x3c736eae30af0cb3ull, 0x3c74a385a63d07a7ull, 0xbc8ff7128fd391f0ull};
    return t[k];
}

// numpy.exp(x) for -700 < x <= 0, float64, as NumPy >= 1.22 computes it on AVX-512 hosts: x = (k + j/16) ln2 + r with
// (k + j/16) = x * log2(e) rounded TOWARD ZERO to a multiple of 1/16 (a fused multiply-add onto 1.5 * 2^48, whose low
// mantissa bits then hold j), r = x - N ln2 with a two-part ln2, a degree-6 polynomial in a fixed FMA order, and
// hi * (p * r + lo) + hi scaled by 2^k.  `tab` = the 32 doubles of np_exp_table_bits (shared or global memory).
YSMR_HD double np_exp_nonpos(double x, const double *tab)
{
    const double shifter = d_from_bits(0x42f8000000003ff0ull);
    const double z = d_fma_rz(x, d_from_bits(0x3ff71547652b82feull), shifter);
    const double n = d_sub(z, shifter);
    const int j = (int)(f64_bits(z) & 15ull);
    double r = d_fma(-n, d_from_bits(0x3fe62e42fefa39efull), x);
    r = d_fma(-n, d_from_bits(0x3c7abc9e3b39803full), r);
    const double r2 = d_mul(r, r);
    double p = d_fma(d_from_bits(0x3f57411836940c04ull), r, d_from_bits(0x3f81101cbbc265c0ull));
    const double p9 = d_fma(d_from_bits(0x3fa55557242d68feull), r, d_from_bits(0x3fc5555553939732ull));
    const double p11 = d_fma(d_from_bits(0x3fe000000000d008ull), r, d_from_bits(0x3fefffffffffff70ull));
    p = d_fma(r2, p, p9);
    p = d_fma(r2, p, p11);
    const double q = d_fma(p, r, tab[16 + j]);
    const double res = d_fma(tab[j], q, tab[j]);
    const int k = (int)floor(n);                                 // n <= 0, a multiple of 1/16
    return d_from_bits(f64_bits(res) + ((unsigned long long)(long long)k << 52));   // res in [1, 2), result normal
}

// gsff.py:179-202: likelihood of one filter, floor 1e-20.
YSMR_HD double gsff_likelihood(double zx, double zy, double ex, double ey, const double *tab)
{
    const double dx = d_sub(zx, ex), dy = d_sub(zy, ey);
    const double e = d_mul(-0.5, d_fma(dy, dy, d_mul(dx, dx)));
    if (!(e > -47.0)) return 1e-20;                              // exp < 4e-21 (also catches NaN): the floor
    const double v = np_exp_nonpos(e, tab);
    return v < 1e-20 ? 1e-20 : v;
}

// One row of numpy.dot(gain_i, flattened history): g0 multiplies the x entries, g1 the y entries of the n newest
// measurements (oldest first), in the dgemv_t order described above.  hist = ring of hist_len (x, y) pairs `stride`
// doubles apart, j0 = index of the oldest of the n.
YSMR_HD double blas_row_dot(const double *g0, const double *g1, bool g1_zero, const double *hist, int64_t stride, int j0,
                            int hist_len, int n)
{
    double l0 = 0.0, l1 = 0.0, l2 = 0.0, l3 = 0.0;
    int j = j0;
    const int nv = n & ~1;
    for (int k = 0; k < nv; k += 2) {
        const double ax = hist[j * stride], ay = g1_zero ? 0.0 : hist[j * stride + 1];
        if (++j == hist_len) j = 0;
        const double bx = hist[j * stride], by = g1_zero ? 0.0 : hist[j * stride + 1];
        if (++j == hist_len) j = 0;
        l0 = d_fma(g0[k], ax, l0); l2 = d_fma(g0[k + 1], bx, l2);
        if (!g1_zero) { l1 = d_fma(g1[k], ay, l1); l3 = d_fma(g1[k + 1], by, l3); }
    }
    double r = d_add(d_add(l0, l2), d_add(l1, l3));
    if (n & 1) {
        const double ax = hist[j * stride], ay = g1_zero ? 0.0 : hist[j * stride + 1];
        r = d_add(r, d_fma(g0[n - 1], ax, g1_zero ? 0.0 : d_mul(g1[n - 1], ay)));
    }
    return r;
}

// ---- GSFF (gsff.py) for one track ----------------------------------------------------------------------------------
// History ring: s.hist[(row * max_tracks + slot) * 2 + {0,1}], row = (frames linked so far) mod hist_len, the same row for
// every track ("ring clock"); a track's hist_n newest rows are valid.

YSMR_HD double *gsff_hist_of(const LinkConfig &c, const LinkState &s, int slot) { return s.hist + 2 * (int64_t)slot; }

// One least-squares FIR estimate (gsff.py:156-177, 230-240) of filter i over the n_i entries ending at row `newest`.
YSMR_HD_NOINLINE void gsff_estimate_one(const LinkConfig &c, const LinkState &s, int slot, int i, int newest)
{
    const double *hist = gsff_hist_of(c, s, slot);
    const int64_t stride = 2 * (int64_t)c.max_tracks;
    const int n = c.n_i[i];
    const double *g = c.gain[i];
    int j = newest + 1 - n; if (j < 0) j += c.hist_len;
    const bool cz = c.cross_zero != 0;
    // row 0 of the gain: (xx, xy) on (x, y); row 1: (yx, yy).  With the reference's gains xy = yx = 0 exactly.
    const double ax = blas_row_dot(g, g + n, cz, hist, stride, j, c.hist_len, n);
    const double ay = cz ? blas_row_dot(g + 3 * n, g + 3 * n, true, hist + 1, stride, j, c.hist_len, n)
                         : blas_row_dot(g + 2 * n, g + 3 * n, false, hist, stride, j, c.hist_len, n);
    s.xh[((int64_t)slot * LINK_MAX_FILTERS + i) * 2] = ax;
    s.xh[((int64_t)slot * LINK_MAX_FILTERS + i) * 2 + 1] = ay;
}

// All active estimates of a track in ONE pass over its history (the general path's form of gsff_estimate_one for the
// reference's gains: no x<->y coupling, identical x and y taps): every entry is loaded once and feeds each filter whose
// horizon reaches back that far; the loads of six entries are issued before their multiply-adds so that their latencies
// overlap.  Per filter and axis the same two FMA chains (even / odd taps, oldest first) as blas_row_dot; an odd horizon's
// newest tap is the separately rounded tail.
YSMR_HD void gsff_estimates_fused(const LinkConfig &c, const LinkState &s, int slot, int mode, int newest)
{
    const double *hist = gsff_hist_of(c, s, slot);
    const int64_t stride = 2 * (int64_t)c.max_tracks;
    const int L = c.hist_len;
    int ni[LINK_MAX_FILTERS];
#pragma unroll
    for (int i = 0; i < LINK_MAX_FILTERS; ++i) ni[i] = i < mode ? c.n_i[i] : 0;
    double ax[LINK_MAX_FILTERS][2], ay[LINK_MAX_FILTERS][2], tx[LINK_MAX_FILTERS], ty[LINK_MAX_FILTERS];
#pragma unroll
    for (int i = 0; i < LINK_MAX_FILTERS; ++i) { ax[i][0] = ax[i][1] = ay[i][0] = ay[i][1] = 0.0; tx[i] = ty[i] = 0.0; }
    const int nmax = ni[mode - 1];
    int e = newest - (nmax - 1); if (e < 0) e += L;
    constexpr int B = 4;
    for (int a0 = nmax - 1; a0 >= 0; a0 -= B) {
        double vx[B], vy[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            int eu = e + u; if (eu >= L) eu -= L;
            const bool on = a0 - u >= 0;
            vx[u] = on ? hist[eu * stride] : 0.0; vy[u] = on ? hist[eu * stride + 1] : 0.0;
        }
        e += B; if (e >= L) e -= L;
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int a = a0 - u;
#pragma unroll
            for (int i = 0; i < LINK_MAX_FILTERS; ++i) {
                if (a >= 0 && a < ni[i]) {
                    const int k = ni[i] - 1 - a;
                    const double g = c.gain[i][k];
                    if (a == 0 && (ni[i] & 1)) { tx[i] = d_mul(g, vx[u]); ty[i] = d_mul(g, vy[u]); }
                    else if (k & 1) { ax[i][1] = d_fma(g, vx[u], ax[i][1]); ay[i][1] = d_fma(g, vy[u], ay[i][1]); }
                    else { ax[i][0] = d_fma(g, vx[u], ax[i][0]); ay[i][0] = d_fma(g, vy[u], ay[i][0]); }
                }
            }
        }
    }
    double *xh = s.xh + (int64_t)slot * LINK_MAX_FILTERS * 2;
#pragma unroll
    for (int i = 0; i < LINK_MAX_FILTERS; ++i) {
        if (i < mode) {
            double rx = d_add(ax[i][0], ax[i][1]), ry = d_add(ay[i][0], ay[i][1]);
            if (ni[i] & 1) { rx = d_add(rx, tx[i]); ry = d_add(ry, ty[i]); }
            xh[2 * i] = rx; xh[2 * i + 1] = ry;
        }
    }
}

YSMR_HD void gsff_estimates(const LinkConfig &c, const LinkState &s, int slot, int mode, int newest)
{
    if (c.cross_zero && c.xy_same && c.hist_len >= 4) gsff_estimates_fused(c, s, slot, mode, newest);
    else for (int i = 0; i < mode; ++i) gsff_estimate_one(c, s, slot, i, newest);
}

// numpy.sum(x_hat_array * weight_array, axis=1) (gsff.py:242, 337): products rounded, summed left to right.
YSMR_HD void gsff_weighted(const double *xh, const double *w, int mode, double *ox, double *oy)
{
    double fx = d_mul(xh[0], w[0]), fy = d_mul(xh[1], w[0]);
    for (int i = 1; i < mode; ++i) { fx = d_add(fx, d_mul(xh[2 * i], w[i])); fy = d_add(fy, d_mul(xh[2 * i + 1], w[i])); }
    *ox = fx; *oy = fy;
}

// correct() then predict() for one track (tracker.py:221-225; gsff.py:251-347, 204-249): (zx, zy) = self.objects[key],
// (ox, oy) = the filtered position that goes to the CSV; the stored position becomes the prediction.  `ring` = the row of
// the current frame.
YSMR_HD void gsff_step(const LinkConfig &c, const LinkState &s, int slot, int ring, double zx, double zy, double *ox, double *oy)
{
    double *hist = gsff_hist_of(c, s, slot);
    const int64_t stride = 2 * (int64_t)c.max_tracks;
    int hist_n = s.hist_n[slot];
    if (hist_n == 0) {                               // first call: previous_measurements = [z] * n_i[0]
        int e = ring - c.n_i[0]; if (e < 0) e += c.hist_len;
        for (int k = 0; k < c.n_i[0]; ++k) { hist[e * stride] = zx; hist[e * stride + 1] = zy; if (++e == c.hist_len) e = 0; }
        hist_n = c.n_i[0];
    }
    int mode = s.mode[slot];
    bool switched = false;
    if (mode < c.n_f) {
        while (hist_n >= c.n_i[mode]) {
            ++mode; switched = true;
            if (mode >= c.n_f) break;
        }
    }
    double *w = s.wgt + (int64_t)slot * LINK_MAX_FILTERS;
    double *xh = s.xh + (int64_t)slot * LINK_MAX_FILTERS * 2;
    if (switched) {
        s.mode[slot] = mode;
        const double w0 = d_div(1.0, (double)mode);  // 1 / mode * np.ones(mode)
        for (int i = 0; i < mode; ++i) w[i] = w0;
        const int prev = ring == 0 ? c.hist_len - 1 : ring - 1;
        gsff_estimates(c, s, slot, mode, prev);
    }
    double p[LINK_MAX_FILTERS];
    double total = 0.0;                              // sum(likelihood_array * weight_array): 0 + p0 + p1 + ...
    for (int i = 0; i < mode; ++i) {
        p[i] = d_mul(gsff_likelihood(zx, zy, xh[2 * i], xh[2 * i + 1], c.exp_tab), w[i]);
        total = d_add(total, p[i]);
    }
    hist[ring * stride] = zx; hist[ring * stride + 1] = zy;
    s.hist_n[slot] = hist_n < c.hist_len ? hist_n + 1 : hist_n;
    for (int i = 0; i < mode; ++i) w[i] = d_div(p[i], total);
    gsff_weighted(xh, w, mode, ox, oy);
    gsff_estimates(c, s, slot, mode, ring);
    gsff_weighted(xh, w, mode, &s.px[slot], &s.py[slot]);
}

// ---- one frame -------------------------------------------------------------------------------------------------------

struct LinkIo {
    const int32_t *blob_count;           // [n_frames]
    const float *blobs;                  // [n_frames][max_blobs][5]
    RowOut *rows;
    long long rows_capacity;
    long long *n_rows;                   // [1] rows written (append == 0) or running total (append == 1)
    int append;                          // 1: start writing at rows[*n_rows] and add to it (chunked pipelines)
    int32_t *status;                     // [1]
    int32_t *first_bad;                  // [1] lowest frame that overflowed, or -1
};

template <class Cta>
YSMR_HD void link_init_track(const LinkConfig &c, const LinkState &s, int slot, int id, const float *det)
{
    s.id[slot] = id;
    s.px[slot] = (double)det[0]; s.py[slot] = (double)det[1];
    s.iw[slot] = det[2]; s.ih[slot] = det[3]; s.ideg[slot] = det[4];
    s.gone[slot] = 0;
    s.mode[slot] = 0; s.hist_n[slot] = 0;
}

// (s2_b, qb) beats (s2_a, qa) under "first index of the minimum ROUNDED distance" (numpy argmin over scipy's cdist row,
// tracker.py:151-163).  Squared distances decide unless they are within 2^-50 relative, where the correctly rounded square
// roots are compared -- so sqrt is almost never evaluated, yet the result is exactly the reference's.
// -1 / 0 / +1 for sqrt(a) <, ==, > sqrt(b) with correctly rounded square roots.  Deliberately not inlined: it is needed
// once in a blue moon and must not be speculated into (or bloat) the hot loops.
YSMR_HD_NOINLINE int sqrt_cmp(double a, double b)
{
    const double da = sqrt(a), db = sqrt(b);
    return da < db ? -1 : (da > db ? 1 : 0);
}

YSMR_HD bool nearer(double s2_a, int qa, double s2_b, int qb)
{
    const double eps = 8.8817841970012523e-16;   // 2^-50
    if (s2_b < s2_a * (1.0 - eps)) return true;
    if (s2_b > s2_a * (1.0 + eps)) return false;
    const int cmp = sqrt_cmp(s2_b, s2_a);
    if (cmp != 0) return cmp < 0;
    return qb < qa;
}

// Uniform grid over the detections of a frame (cells of `cell` pixels, gw x gh of them, coordinates clamped into it):
// the exact nearest detection of a point is found by visiting the rings of cells around the point's cell until the best
// distance found is provably smaller than the distance to anything not visited yet.  A clamped detection really lies
// further out than its cell says, so the bound stays valid; a side of the visited square that has reached the edge of the
// grid has nothing beyond it.
struct DetGrid {
    const float2 *dxy;                   // [m] detection centres
    const uint32_t *cell_start;          // [gw*gh + 1]
    const int32_t *cell_items;           // [m] detection indices grouped by cell
    int gw, gh;
    double cell, inv_cell;
};

YSMR_HD int grid_coord(double v, double inv_cell, int g)
{
    double t = floor(v * inv_cell);
    t = fmax(t, 0.0);                    // (NaN -> 0)
    t = fmin(t, (double)(g - 1));
    return (int)t;
}

YSMR_HD void grid_visit(const DetGrid &G, int cx, int cy, double ox, double oy, double &best, int &arg)
{
    const int cidx = cy * G.gw + cx;
    const uint32_t a = G.cell_start[cidx], b = G.cell_start[cidx + 1];
    for (uint32_t t = a; t < b; ++t) {
        const int q = G.cell_items[t];
        const float2 d = G.dxy[q];
        const double dx = d_sub(ox, (double)d.x), dy = d_sub(oy, (double)d.y);
        const double s2 = d_add(d_mul(dx, dx), d_mul(dy, dy));      // scipy euclidean: s += d*d per coordinate
        if (arg == 0x7fffffff || nearer(best, arg, s2, q)) { best = s2; arg = q; }
    }
}

// first index of the minimum rounded distance over all m > 0 detections; *s2_out = its squared distance
YSMR_HD int grid_nearest(const DetGrid &G, double ox, double oy, double *s2_out)
{
    const int cx = grid_coord(ox, G.inv_cell, G.gw), cy = grid_coord(oy, G.inv_cell, G.gh);
    double best = 1.0e300; int arg = 0x7fffffff;
    const int kmax = (G.gw > G.gh ? G.gw : G.gh);
    for (int k = 0; k <= kmax; ++k) {
        const int x0 = cx - k, x1 = cx + k, y0 = cy - k, y1 = cy + k;
        // the ring of cells at Chebyshev distance k, clipped to the grid (one visit site: rows y0 and y1 in full, the two
        // end cells of the rows between)
        const int xa = x0 < 0 ? 0 : x0, xb = x1 >= G.gw ? G.gw - 1 : x1;
        const int ya = y0 < 0 ? 0 : y0, yb = y1 >= G.gh ? G.gh - 1 : y1;
        for (int y = ya; y <= yb; ++y) {
            const bool full = y == y0 || y == y1;
            const int step = full || x1 == x0 ? 1 : x1 - x0;             // interior rows: only x0 and x1
            for (int xx = full ? xa : x0; xx <= xb; xx += step)
                if (xx >= 0) grid_visit(G, xx, y, ox, oy, best, arg);
        }
        // everything not visited yet lies outside the square of cells [x0, x1] x [y0, y1]
        const bool L = x0 <= 0, R = x1 >= G.gw - 1, T = y0 <= 0, B = y1 >= G.gh - 1;
        if (L && R && T && B) break;
        if (arg != 0x7fffffff) {
            double bound = 1.0e300;
            if (!L) bound = fmin(bound, ox - (double)x0 * G.cell);
            if (!R) bound = fmin(bound, (double)(x1 + 1) * G.cell - ox);
            if (!T) bound = fmin(bound, oy - (double)y0 * G.cell);
            if (!B) bound = fmin(bound, (double)(y1 + 1) * G.cell - oy);
            // strictly nearer than anything outside, with room for the float32 coordinates and the roundings
            if (bound > 0.0 && best < bound * bound * (1.0 - 1.0e-9)) break;
        }
    }
    *s2_out = best;
    return arg;
}

// Per-frame scratch of the general path: shared memory on the device when it fits, global memory otherwise; plain arrays in
// the host emulation.
struct FrameScratch {
    float2 *dxy;                         // [max_blobs]
    unsigned long long *col_best;        // [max_blobs] bits of the smallest row-minimum that chose this detection
    int32_t *col_row;                    // [max_blobs] winning row (and scratch of the counting sort)
    int32_t *cell_items;                 // [max_blobs]
    uint32_t *cell_start;                // [LINK_GRID_CELLS + 2]
    int32_t *flags;                      // [4]: 0 = some detection has two claimants at the same distance
};

constexpr int LINK_GRID_CELLS = 1024;    // at most 32 x 32 cells

// Staging of a frame's detections for the general path: centres into f.dxy, claim slots reset, detections binned into the
// uniform grid by a counting sort over cells.  On return G points at the grid and every thread may search it.
template <class Cta>
YSMR_HD void stage_detections_generic(Cta &cta, const LinkConfig &c, const LinkScratch &x, const FrameScratch &f, const float *dets,
                                      int m, DetGrid &G)
{
    const int tid = cta.tid(), nthr = cta.nthr();
    const int NONE = 0x7fffffff;
    G.dxy = f.dxy; G.cell_start = f.cell_start; G.cell_items = f.cell_items;
    const int ncell = G.gw * G.gh;
    for (int k = tid; k <= ncell; k += nthr) f.cell_start[k] = 0;
    if (tid == 0) f.flags[0] = 0;
    cta.sync();
    for (int q = tid; q < m; q += nthr) {
        float2 d; d.x = dets[5 * q]; d.y = dets[5 * q + 1];
        f.dxy[q] = d;
        f.col_best[q] = ~0ull;
        const int cell = grid_coord((double)d.y, G.inv_cell, G.gh) * G.gw + grid_coord((double)d.x, G.inv_cell, G.gw);
        x.flag[q] = (uint32_t)cell;
        f.col_row[q] = (int32_t)cta.atomic_add_u32(&f.cell_start[cell], 1u);   // position within the cell
    }
    cta.sync();
    cta.exclusive_scan(f.cell_start, ncell + 1);
    for (int q = tid; q < m; q += nthr) {
        f.cell_items[f.cell_start[x.flag[q]] + (uint32_t)f.col_row[q]] = q;
        f.col_row[q] = NONE;
    }
    cta.sync();
}

// Processes frames [start_frame, n_frames) of the chunk.  One CTA; within a frame the phases are loops over tracks or
// detections separated by cta.sync().  Header values live in registers of every thread and are updated identically by all
// of them (every quantity they depend on is CTA-uniform), thread 0 writes them back at the end.  Track state is
// struct-of-arrays in global memory, the history ring entry-major ([row][slot], row = frame counter mod hist_len), so that
// threads working on neighbouring slots touch neighbouring addresses.
template <class Cta>
YSMR_HD void link_chunk(Cta &cta, const LinkConfig &c, const LinkState &s, const LinkScratch &x, const FrameScratch &f,
                        const LinkIo &io, int first_frame, int n_frames, int start_frame = 0)
{
    const int tid = cta.tid(), nthr = cta.nthr();
    int n = s.hdr[0], next_id = s.hdr[1], n_free = s.hdr[2], sel = s.hdr[3];
    const int clock0 = s.hdr[4] - start_frame;   // frames linked before this launch: the ring clock (the fast path has
                                                 // already counted the start_frame frames it handled)
    long long rows_total = io.append ? *io.n_rows : 0;
    bool row_overflow = false;
    const int NONE = 0x7fffffff;

    for (int fi = start_frame; fi < n_frames; ++fi) {
        const int m = io.blob_count[fi];
        const float *dets = io.blobs + (int64_t)fi * c.max_blobs * 5;
        int32_t *order = s.order[sel];
        bool age_unmatched = false;      // rows with flag[r] == 0 are aged afterwards
        int births = 0;

        if (m == 0) {
            for (int r = tid; r < n; r += nthr) x.flag[r] = 0;              // nobody matched
            age_unmatched = true;
            cta.sync();
        } else if (n == 0) {
            births = m;
            for (int q = tid; q < m; q += nthr) x.list[q] = q;              // detection order (tracker.py:135-137)
            cta.sync();
        } else {
            // ---- stage the detections and bin them into the grid (counting sort over cells): the Cta policy decides where
            // the grid lives (stage_detections_generic below; a cooperative grid builds a private copy per block)
            DetGrid G;
            G.cell = c.grid_cell; G.inv_cell = 1.0 / c.grid_cell; G.gw = c.grid_w; G.gh = c.grid_h;
            cta.stage_detections(c, x, f, dets, m, G);
            // ---- nearest detection of every track (cdist row minimum / first argmin, tracker.py:151-163) and its claim
            for (int r = tid; r < n; r += nthr) {
                const int slot = order[r];
                double s2;
                const int q = grid_nearest(G, s.px[slot], s.py[slot], &s2);
                const double d = sqrt(s2);
                x.row_min[r] = d; x.row_arg[r] = q;
                if (c.max_distance <= 0.0 || d <= c.max_distance)
                    if (cta.atomic_min_u64(&f.col_best[q], f64_bits(d)) == f64_bits(d)) f.flags[0] = 1;   // same distance twice
            }
            cta.sync();
            // the smallest rounded distance wins a detection, the lowest row among equal distances (tracker.py:158-189)
            const bool ties = f.flags[0] != 0;
            if (ties) {
                for (int r = tid; r < n; r += nthr) {
                    const int q = x.row_arg[r];
                    if (f.col_best[q] == f64_bits(x.row_min[r]) && (c.max_distance <= 0.0 || x.row_min[r] <= c.max_distance))
                        cta.atomic_min_i32(&f.col_row[q], r);
                }
                cta.sync();
            }
            for (int r = tid; r < n; r += nthr) {
                const int q = x.row_arg[r];
                const bool ok = c.max_distance <= 0.0 || x.row_min[r] <= c.max_distance;
                const bool won = ties ? f.col_row[q] == r : (ok && f.col_best[q] == f64_bits(x.row_min[r]));
                x.flag[r] = won ? 1u : 0u;
                if (won) {                                                  // tracker.py:181-184
                    const int slot = order[r];
                    const float *d = dets + 5 * q;
                    s.px[slot] = (double)d[0]; s.py[slot] = (double)d[1];
                    s.iw[slot] = d[2]; s.ih[slot] = d[3]; s.ideg[slot] = d[4];
                    s.gone[slot] = 0;
                    if (!ties) f.col_row[q] = r;                            // marks the detection as used
                }
            }
            cta.sync();
            if (n >= m) {
                age_unmatched = true;                                       // tracker.py:198-211
            } else {                                                        // tracker.py:215-217
                for (int q = tid; q <= m; q += nthr) x.flag[q] = (q < m && f.col_row[q] == NONE) ? 1u : 0u;
                cta.sync();
                births = (int)cta.exclusive_scan(x.flag, m + 1);
                for (int q = tid; q < m; q += nthr)
                    if (x.flag[q + 1] != x.flag[q]) x.list[x.flag[q]] = q;  // unused detections, ascending
                cta.sync();
                if (tid == 0) cpython_set_order(x.list, births, x.table);
                cta.sync();
            }
        }

        if (age_unmatched) {
            // flag[r] = 1 matched.  Age the others, drop those beyond max_disappeared, compact `order`.
            int drop = 0;
            for (int r = tid; r < n; r += nthr) {
                uint32_t keep = 1;
                if (!x.flag[r]) {
                    const int slot = order[r];
                    const int g = s.gone[slot] + 1;
                    s.gone[slot] = g;
                    s.iw[slot] = 0.f; s.ih[slot] = 0.f; s.ideg[slot] = 0.f;   // [0] * len(info)
                    if ((double)g > c.max_disappeared) { keep = 0; drop = 1; }
                }
                x.flag[r] = keep;
            }
            if (tid == 0) x.flag[n] = 0;
            if (cta.any(drop)) {                                            // (barrier + vote)
                const int kept = (int)cta.exclusive_scan(x.flag, n + 1);
                int32_t *order2 = s.order[sel ^ 1];
                for (int r = tid; r < n; r += nthr) {
                    const int before = (int)x.flag[r];
                    if ((int)x.flag[r + 1] != before) order2[before] = order[r];
                    else s.free_slots[n_free + (r - before)] = order[r];
                }
                cta.sync();
                n_free += n - kept;
                n = kept;
                sel ^= 1;
                order = s.order[sel];
            }
        }
        if (births > 0) {
            int room = c.max_tracks - n;
            if (births > room) {
                if (tid == 0) { cta.atomic_or_i32(io.status, LINK_ST_TRACK_OVERFLOW); cta.atomic_min_i32(io.first_bad, first_frame + fi); }
                births = room;
            }
            for (int b = tid; b < births; b += nthr) {
                const int slot = s.free_slots[n_free - 1 - b];
                order[n + b] = slot;
                link_init_track<Cta>(c, s, slot, next_id + b, dets + 5 * x.list[b]);
            }
            cta.sync();
            n += births; next_id += births; n_free -= births;
        }

        // filter + emit (tracker.py:219-227, track_eval.py:313-316)
        const bool room = rows_total + n <= io.rows_capacity;
        int ring = (clock0 + fi) % c.hist_len; if (ring < 0) ring += c.hist_len;
        for (int r = tid; r < n; r += nthr) {
            const int slot = order[r];
            double ox = s.px[slot], oy = s.py[slot];
            if (c.use_gsff) gsff_step(c, s, slot, ring, ox, oy, &ox, &oy);
            if (room) {
                RowOut &o = io.rows[rows_total + r];
                o.frame = first_frame + fi; o.track_id = s.id[slot];
                o.x = ox; o.y = oy; o.w = s.iw[slot]; o.h = s.ih[slot]; o.deg = s.ideg[slot]; o.pad = 0;
            }
        }
        if (room) rows_total += n;
        else if (!row_overflow) {
            row_overflow = true;
            if (tid == 0) { cta.atomic_or_i32(io.status, LINK_ST_ROW_OVERFLOW); cta.atomic_min_i32(io.first_bad, first_frame + fi); }
        }
        cta.sync();
    }
    if (tid == 0) {
        s.hdr[0] = n; s.hdr[1] = next_id; s.hdr[2] = n_free; s.hdr[3] = sel;
        s.hdr[4] = clock0 + (n_frames > start_frame ? n_frames : start_frame); s.hdr[5] = n;
        *io.n_rows = rows_total;
    }
}

#if defined(__CUDACC__)
// Gate of a pipelined launch (capi.cu: track_chunks): the sequential kernel is enqueued BEFORE the detections it will
// read exist and spins on *ready (set by a memset behind the detection kernels of the chunk), so that it keeps the SM it
// runs on from chunk to chunk instead of having to wait, every chunk, for an SM that the detection kernels of the next
// chunk have completely vacated.  The candidate tables are then built on the detection stream (launch_link_prep), into half
// `table` of the double-buffered table memory.  The caller enqueues, in this order: detection, launch_link_prep, the memset
// of the flag, and only then launch_link -- so that even a tool that serialises kernels (ncu) never runs the waiting kernel
// before the work it waits for has been submitted.
struct LinkGate {
    const int32_t *ready;
    int table;
};
// the candidate tables of a gated launch, on the detection stream (before the flag is set; launch_link then skips them)
cudaError_t launch_link_prep(const LinkConfig &c, const LinkScratch &x, const int32_t *blob_count, const float *blobs, int n_frames,
                             int table, cudaStream_t st);
cudaError_t launch_link(const LinkConfig &c, const LinkState &s, const LinkScratch &x, const FrameScratch &f, const LinkIo &io,
                        int first_frame, int n_frames, int allow_fast, cudaStream_t st, const LinkGate *gate = nullptr);
cudaError_t launch_link_reset(const LinkState &s, int max_tracks, cudaStream_t st);
cudaError_t link_kernel_init();            // per device: shared-memory opt-in of the linker kernels (from ysmr_create)
#endif

}  // namespace ysmr
