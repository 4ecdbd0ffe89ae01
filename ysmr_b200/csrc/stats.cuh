// Per-track statistics of evaluate_tracks() (partial): the columns of df_stats that are reductions over
// a track's rows -- /root/reference/ysmr/track_eval.py:905-945 (deltas, travelled distance, `moving` and its two median
// filters), :1030-1062 (time, distance, displacement = largest pairwise distance, percent motile, speed, arc-chord ratio,
// bacteria length).  Turn points, motility phenotype and median speed (:946-1029) are not built.
//
// Exactness: pandas' groupby sum / mean are Kahan-compensated sequential sums (pandas/_libs/groupby.pyx group_sum /
// group_mean), so those two are restated as such (kahan_*); `bac_length` is a float16 column whose group mean pandas computes
// in float32; everything else is element-wise IEEE arithmetic or order-independent (max, integer sums).
#pragma once
#include "common.cuh"

#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

namespace ysmr {

struct StatCols { const uint32_t *t; const double *x, *y, *w, *h; };
struct StatCfg { double px, fps; int kernel2; };          // 'pixel per micrometre', fps, the second median-filter size (:933-936)
enum { STAT_DISTANCE = 0, STAT_SPEED, STAT_TIME, STAT_DISPLACEMENT, STAT_PERC_MOTILE, STAT_ACR, STAT_BAC_LENGTH, STAT_DISPL_BY_LENGTH,
       STAT_COLUMNS };

// the float32 value of numpy's float64 -> float16 conversion (round to nearest even) of a finite v >= 0
YSMR_HD float f16_round(double v)
{
    if (!(v > 0.0)) return 0.f;
    if (v >= 65520.0) return INFINITY;
    int e = ilogb(v);
    if (e < -14) e = -14;                                   // subnormal halves share the spacing 2^-24
    const double ulp = ldexp(1.0, e - 10);
    return (float)(rint(v / ulp) * ulp);
}

// row i of a track that starts at row lo: travelled_dist (:926) and the raw `moving` flag (:927-929)
YSMR_HD double travelled_dist(const StatCols &c, double px, int lo, int i)
{
    if (i == lo) return 0.0;                                // x_delta = y_delta = 0 at a track start (:909)
    const double dx = c.x[i] - c.x[i - 1], dy = c.y[i] - c.y[i - 1];
    return sqrt(dx * dx + dy * dy) / px;
}
YSMR_HD int moving_raw(const StatCols &c, double px, int lo, int i)
{
    const double dt = i == lo ? 1.0 : (double)c.t[i] - (double)c.t[i - 1];      // t_delta = 1 at a track start (:910)
    return travelled_dist(c, px, lo, i) / dt > 1.0e-3 ? 1 : 0;
}
// scipy.signal.medfilt of a 0/1 sequence with zero padding: 1 iff at least (k + 1) / 2 ones in the window
YSMR_HD int medfilt_bit(const uint8_t *m, int lo, int hi, int i, int k)
{
    int ones = 0;
    for (int j = i - k / 2; j <= i + k / 2; ++j) ones += (j >= lo && j <= hi) ? m[j] : 0;
    return ones >= (k + 1) / 2 ? 1 : 0;
}
// pandas group_sum over float64 (Kahan), rows lo..hi of travelled_dist
YSMR_HD double kahan_distance(const StatCols &c, double px, int lo, int hi)
{
    double s = 0.0, comp = 0.0;
    for (int i = lo; i <= hi; ++i) {
        const double y = travelled_dist(c, px, lo, i) - comp;
        const double t = s + y;
        comp = (t - s) - y;
        s = t;
    }
    return s;
}
// pandas group_mean of the float16 column bac_length (:923), computed in float32 (Kahan), divided by the row count
YSMR_HD float kahan_bac_length(const StatCols &c, double px, int lo, int hi)
{
    float s = 0.f, comp = 0.f;
    for (int i = lo; i <= hi; ++i) {
        const double w = c.w[i] / px, h = c.h[i] / px;
        const float v = f16_round(w >= h ? w : h);
        const float y = v - comp;
        const float t = s + y;
        comp = (t - s) - y;
        s = t;
    }
    return s / (float)(hi - lo + 1);
}
// the scalar tail (:1043-1096) once the reductions are known
YSMR_HD void finish_statistics(const StatCols &c, const StatCfg &g, int lo, int hi, double dist, double max_d2, long long motile_total,
                               float bac_length, double *out)
{
    const long long t_last = (long long)c.t[hi] - (long long)c.t[lo];          // t_norm of the last row
    const double time_s = (double)(t_last + 1) / g.fps;
    const double xn = (c.x[hi] - c.x[lo]) / g.px, yn = (c.y[hi] - c.y[lo]) / g.px;
    const double chord = sqrt(xn * xn + yn * yn);
    const double disp = sqrt(max_d2);
    out[STAT_DISTANCE] = dist;
    out[STAT_SPEED] = motile_total != 0 ? dist / time_s : 0.0;
    out[STAT_TIME] = time_s;
    out[STAT_DISPLACEMENT] = disp;
    out[STAT_PERC_MOTILE] = (double)motile_total / (double)(t_last + 1) * 100.0;
    out[STAT_ACR] = dist != 0.0 ? chord / dist : 0.0;
    out[STAT_BAC_LENGTH] = (double)bac_length;
    out[STAT_DISPL_BY_LENGTH] = bac_length != 0.f ? disp / (double)bac_length : 0.0;
}

// everything for one track by one thread (host emulation; the device kernel in select.cu spreads the loops over a CTA)
YSMR_HD void track_statistics_serial(const StatCols &c, const StatCfg &g, int lo, int hi, uint8_t *ma, uint8_t *mb, double *out)
{
    for (int i = lo; i <= hi; ++i) ma[i] = (uint8_t)moving_raw(c, g.px, lo, i);
    for (int i = lo; i <= hi; ++i) mb[i] = (uint8_t)medfilt_bit(ma, lo, hi, i, 3);
    long long motile = 0;
    for (int i = lo; i <= hi; ++i) motile += medfilt_bit(mb, lo, hi, i, g.kernel2);
    double max_d2 = 0.0;
    for (int i = lo; i <= hi; ++i) {
        const double xi = (c.x[i] - c.x[lo]) / g.px, yi = (c.y[i] - c.y[lo]) / g.px;
        for (int j = i + 1; j <= hi; ++j) {
            const double dx = xi - (c.x[j] - c.x[lo]) / g.px, dy = yi - (c.y[j] - c.y[lo]) / g.px;
            const double d2 = dx * dx + dy * dy;
            max_d2 = d2 > max_d2 ? d2 : max_d2;
        }
    }
    finish_statistics(c, g, lo, hi, kahan_distance(c, g.px, lo, hi), max_d2, motile, kahan_bac_length(c, g.px, lo, hi), out);
}

}  // namespace ysmr
