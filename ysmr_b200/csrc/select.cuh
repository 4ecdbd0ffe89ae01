// Track selection: the per-track logic of find_good_tracks(), /root/reference/ysmr/track_eval.py:408-538,
// written once as YSMR_HD code over a small "lanes" abstraction: on the device one warp walks one track (select.cu), the
// host emulation (tests/host_emul, test infrastructure) runs the same source with a single lane.
//
// What has to be reproduced exactly, not just "statistically": the reference decides with comparisons of pandas means
// against bounds that are themselves data quantiles, so the mean must be pandas' mean bit for bit --
// Series.mean() = ndarray.sum() / count, and ndarray.sum() of a contiguous float64 array is numpy's pairwise summation
// over all n values (DOUBLE_add's reduce loop, numpy/_core/src/umath/loops_utils.h.src; checked bit for bit against
// numpy 2.3 and pandas 3.0 in tests/test_select_host.py).  np_pairwise_sum below is that loop.  Everything order-independent (largest
// frame gap and where it first occurs, first motility outlier, min / max) is a plain lane-parallel reduction.
#pragma once
#include "common.cuh"

namespace ysmr {

// numpy's @TYPE@_pairwise_sum for unit-stride float64 (PW_BLOCKSIZE 128, eight accumulators): blocks of at most 128 values ...
YSMR_HD double np_pairwise_block(const double *a, int64_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
        r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += a[i];
    return res;
}
// ... and above that the recursion  sum(a, n) = sum(a, n2) + sum(a + n2, n - n2),  n2 = n / 2 rounded down to a multiple of 8,
// evaluated here with an explicit stack (no device-side recursion: the stack size stays static)
YSMR_HD_NOINLINE double np_pairwise_sum(const double *a, int64_t n)
{
    struct Frame { const double *a; int64_t n; int phase; };
    Frame fr[40];
    double vals[40];
    int fp = 0, vp = 0;
    fr[fp].a = a; fr[fp].n = n; fr[fp].phase = 0; ++fp;
    while (fp > 0) {
        Frame &f = fr[fp - 1];
        if (f.n <= 128) { vals[vp++] = np_pairwise_block(f.a, f.n); --fp; continue; }
        int64_t n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.phase == 0) { f.phase = 1; fr[fp].a = f.a; fr[fp].n = n2; fr[fp].phase = 0; ++fp; }
        else if (f.phase == 1) { f.phase = 2; fr[fp].a = f.a + n2; fr[fp].n = f.n - n2; fr[fp].phase = 0; ++fp; }
        else { const double right = vals[--vp], left = vals[--vp]; vals[vp++] = left + right; --fp; }
    }
    return vals[0];
}
// ndarray.sum() / count for n >= 1 contiguous float64 values (pandas nanops.nanmean without missing values)
YSMR_HD double np_mean(const double *a, int64_t n)
{
    return np_pairwise_sum(a, n) / (double)n;
}

// Columns of the data frame after the initial clean-up (track_eval.py:609-698, 716-741), rows grouped by track.
struct SelectCols {
    const uint32_t *t;        // POSITION_T
    const double *x, *y;      // POSITION_X, POSITION_Y
    const double *area;       // WIDTH * HEIGHT
    const double *ratio;      // ratio_wh (track_eval.py:698)
    const int8_t *outlier;    // df['distance'] after the outer-fence test (track_eval.py:721)
};
struct SelectCfg {
    int min_len;              // minimal_length_frames
    int max_holes;            // 'maximal consecutive holes'
    int max_recursion;        // 'maximal recursion depth'
    double max_empty;         // 'maximal empty frames in %' (already / 100 + 1)
    double lower, upper;      // area bounds: the quantiles, or -1 / +inf
    double ratio_min, ratio_max;
    double edge;              // 'percent of screen edges to exclude' (fraction)
    int frame_h, frame_w;
    int limit_frames, limit_exactly;
};
struct SelectSeg { int start, stop, depth; };

// Single-lane execution (host emulation; also what a lone thread would do)
struct OneLane {
    YSMR_HD int lane() const { return 0; }
    YSMR_HD int lanes() const { return 1; }
    YSMR_HD void max_first(double &, int &) const {}
    YSMR_HD int sum(int v) const { return v; }
    YSMR_HD int min_i(int v) const { return v; }
    YSMR_HD double min_d(double v) const { return v; }
    YSMR_HD double max_d(double v) const { return v; }
    YSMR_HD double bcast(double v) const { return v; }
};

// find_good_tracks for the rows [start, stop] of one track, with the reference's recursion as an explicit depth-first
// stack (stack_cap >= max_recursion + 2).  Returns the kick reason (track_eval.py:441-452); *good_start / *good_stop = the
// longest accepted fragment (the first among equals, track_eval.py:771-779) after the length limit (:781-794), or -1.
template <class W>
YSMR_HD int select_track(const W &wp, const SelectCols &c, const SelectCfg &g, int start, int stop, SelectSeg *stack, int stack_cap,
                         int *good_start, int *good_stop)
{
    int best_s = -1, best_e = -1, best_len = 0, kick_min = 8, sp = 0;
    if (wp.lane() == 0) { stack[0].start = start; stack[0].stop = stop; stack[0].depth = 0; }
    sp = 1;
    const int lane = wp.lane(), lanes = wp.lanes();
    while (sp > 0) {
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        const SelectSeg seg = stack[--sp];
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        const int s = seg.start, e = seg.stop, size = e - s + 1;
        int kick = 8;
        bool split = false; int a_s = 0, a_e = -1, b_s = 0, b_e = -1;
        if (size >= g.min_len) {
            kick = 7;
            // largest gap of POSITION_T.diff() and the first row where it occurs (track_eval.py:461-462, 505)
            double md = -1.0; int mi = 0x7fffffff;
            for (int i = s + 1 + lane; i <= e; i += lanes) {
                const double d = (double)c.t[i] - (double)c.t[i - 1];
                if (d > md) { md = d; mi = i; }
            }
            wp.max_first(md, mi);
            if (size >= 2 && md <= (double)g.max_holes) {
                kick = 6;
                int cnt = 0, first = 0x7fffffff;
                for (int i = s + lane; i <= e; i += lanes)
                    if (c.outlier[i]) { ++cnt; first = first < i ? first : i; }
                cnt = wp.sum(cnt); first = wp.min_i(first);
                if (cnt == 0) {
                    kick = 5;
                    const double duration = (double)(uint32_t)(c.t[e] - c.t[s] + 1u);
                    if (duration / (double)size < g.max_empty) {
                        kick = 4;
                        double m = 0.0;
                        if (lane == 0) m = np_mean(c.area + s, size);
                        m = wp.bcast(m);
                        if (g.lower <= m && m <= g.upper) {
                            kick = 3;
                            if (lane == 0) m = np_mean(c.ratio + s, size);
                            m = wp.bcast(m);
                            if (g.ratio_min < m && m < g.ratio_max) {
                                kick = 2;
                                if (lane == 0) m = np_mean(c.y + s, size);
                                m = wp.bcast(m);
                                if (g.edge * (double)g.frame_h < m && m < (1.0 - g.edge) * (double)g.frame_h) {
                                    if (lane == 0) m = np_mean(c.x + s, size);
                                    m = wp.bcast(m);
                                    if (g.edge * (double)g.frame_w < m && m < (1.0 - g.edge) * (double)g.frame_w) {
                                        kick = 1;
                                        double xmin = 1.0e300, xmax = -1.0e300, ymin = 1.0e300, ymax = -1.0e300;
                                        for (int i = s + lane; i <= e; i += lanes) {
                                            const double xv = c.x[i], yv = c.y[i];
                                            xmin = xv < xmin ? xv : xmin; xmax = xv > xmax ? xv : xmax;
                                            ymin = yv < ymin ? yv : ymin; ymax = yv > ymax ? yv : ymax;
                                        }
                                        xmin = wp.min_d(xmin); xmax = wp.max_d(xmax); ymin = wp.min_d(ymin); ymax = wp.max_d(ymax);
                                        if (g.edge == 0.0 || !(xmin < 0.0 || xmax > (double)g.frame_w || ymin < 0.0 || ymax > (double)g.frame_h)) {
                                            kick = 0;
                                            if (size > best_len) { best_len = size; best_s = s; best_e = e; }
                                        }
                                    }
                                }
                            }
                        }
                    }
                } else {          // split at the first outlier, which is left out (track_eval.py:498-501)
                    split = true; a_s = s; a_e = first - 1; b_s = first + 1; b_e = e;
                }
            } else if (size >= 2) {   // split before the largest gap (track_eval.py:503-506)
                split = true; a_s = s; a_e = mi - 1; b_s = mi; b_e = e;
            }
        }
        kick_min = kick < kick_min ? kick : kick_min;
        if (split && seg.depth < g.max_recursion) {
            const int need = g.min_len < 3 ? 3 : g.min_len;               // track_eval.py:514-519
            const bool take_a = a_e - a_s + 1 >= need, take_b = b_e - b_s + 1 >= need;
            if (sp + 2 <= stack_cap) {
                if (lane == 0) {
                    int q = sp;
                    if (take_b) { stack[q].start = b_s; stack[q].stop = b_e; stack[q].depth = seg.depth + 1; ++q; }
                    if (take_a) { stack[q].start = a_s; stack[q].stop = a_e; stack[q].depth = seg.depth + 1; ++q; }
                }
                sp += (take_a ? 1 : 0) + (take_b ? 1 : 0);
            }
        }
    }
    // length limit (track_eval.py:781-794)
    if (best_s >= 0 && g.limit_frames) {
        const uint32_t limit = (uint32_t)g.limit_frames + c.t[best_s] - 1u;
        int last = -1;
        for (int i = best_s + lane; i <= best_e; i += lanes) {
            const bool hit = g.limit_exactly ? c.t[i] == limit : c.t[i] <= limit;
            if (hit) last = i;                                           // POSITION_T increases: the last hit is idxmax
        }
        last = -wp.min_i(-last);
        if (last < 0) best_s = -1;
        best_e = last;
    }
    *good_start = best_s; *good_stop = best_s >= 0 ? best_e : -1;
    return kick_min;
}

}  // namespace ysmr
