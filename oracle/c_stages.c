/* ORACLE (test infrastructure only; never linked into libysmr_b200.so).
 *
 * Plain-C restatement of the integer / float32 arithmetic of the front-end stages of YSMR's hot loop.  The
 * reference calls un-vendored OpenCV (4.13.0 in this image) for them; the published/observed arithmetic is
 * restated here and pinned by tests/test_oracle_stages.py against cv2 itself on the reference's call sites:
 *
 *   track_eval.py:180      cv2.cvtColor(BGR2GRAY)        -> ysmr_oracle_grey
 *   track_eval.py:182      cv2.GaussianBlur((3,3), 0)    -> ysmr_oracle_blur3
 *   track_eval.py:189-208  cv2.adaptiveThreshold(GAUSSIAN_C, 11) mean image + compare
 *                                                         -> ysmr_oracle_gauss11_mean, ysmr_oracle_compare
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).  -ffp-contract=off matters: every FMA
 * below is explicit (fmaf) and every other float operation must round separately.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* OpenCV's fixed-point luma, Q15 with rounding. */
void ysmr_oracle_grey(const uint8_t *bgr, uint8_t *grey, int64_t n_px)
{
    for (int64_t i = 0; i < n_px; ++i) {
        uint32_t b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        grey[i] = (uint8_t)((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15);
    }
}

static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}

static inline int clampi(int p, int n) { return p < 0 ? 0 : (p >= n ? n - 1 : p); }

/* 3x3 binomial blur, BORDER_REFLECT_101, (sum + 8) >> 4, all integer. */
void ysmr_oracle_blur3(const uint8_t *src, uint8_t *dst, int h, int w)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t *r0 = src + (int64_t)reflect101(y - 1, h) * w;
        const uint8_t *r1 = src + (int64_t)y * w;
        const uint8_t *r2 = src + (int64_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; ++x) {
            int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
            int s = r0[xl] + 2 * r0[x] + r0[xr] + 2 * r1[xl] + 4 * r1[x] + 2 * r1[xr] + r2[xl] + 2 * r2[x] + r2[xr];
            dst[(int64_t)y * w + x] = (uint8_t)((s + 8) >> 4);
        }
    }
}

/* SURVEY A.3 (extended by probing cv2 4.13.0 for every width 1..80 and 1224..1232):
 * float32 separable 11-tap Gaussian exactly as OpenCV's AVX2 filter path evaluates it.
 *   k[11]      the taps, as returned by cv2.getGaussianKernel(11, 0, CV_32F) (passed in by the caller)
 *   row pass   acc = k0*x0 ; acc = fma(k_i, x_i, acc), i = 1..10, left to right; BORDER_REPLICATE.
 *              The last w%4 columns are OpenCV's scalar tail: taps 1..8 round the multiply and the add
 *              separately, taps 9 and 10 are fused.
 *   col pass   acc = k5*r5 ; s = r[5-j] + r[5+j] ; acc = fma(k[5+j], s, acc), j = 1..5.
 *              The last w%8 columns (scalar tail of the 8-wide vector loop) round multiply and add separately.
 * mean_f receives the float image, mean_u8 its cvRound (rint, half to even) saturated to u8. */
void ysmr_oracle_gauss11_mean(const uint8_t *src, const float *k, float *mean_f, uint8_t *mean_u8, int h, int w)
{
    const int row_tail_from = w - (w % 4);
    const int col_tail_from = w - (w % 8);
    float *rowbuf = (float *)malloc(sizeof(float) * (size_t)h * (size_t)w);
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (int64_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = k[0] * (float)s[clampi(x - 5, w)];
            for (int i = 1; i < 11; ++i) {
                float v = (float)s[clampi(x - 5 + i, w)];
                if (x < row_tail_from || i >= 9) acc = fmaf(k[i], v, acc);
                else { float p = k[i] * v; acc = acc + p; }
            }
            rowbuf[(int64_t)y * w + x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            float acc = k[5] * rowbuf[(int64_t)y * w + x];
            for (int j = 1; j <= 5; ++j) {
                float a = rowbuf[(int64_t)clampi(y - j, h) * w + x];
                float b = rowbuf[(int64_t)clampi(y + j, h) * w + x];
                float s = a + b;
                if (x < col_tail_from) acc = fmaf(k[5 + j], s, acc);
                else { float p = k[5 + j] * s; acc = acc + p; }
            }
            if (mean_f) mean_f[(int64_t)y * w + x] = acc;
            if (mean_u8) {
                float r = rintf(acc);
                mean_u8[(int64_t)y * w + x] = (uint8_t)(r < 0.f ? 0 : (r > 255.f ? 255 : (int)r));
            }
        }
    }
    free(rowbuf);
}

/* adaptiveThreshold's table compare: d = src - mean; BINARY: 255 if d > t ; BINARY_INV: 255 if d <= t.
 * (t = -ceil(C) resp. -floor(C), computed by the caller exactly as OpenCV does.) */
void ysmr_oracle_compare(const uint8_t *src, const uint8_t *mean, uint8_t *dst, int64_t n_px, int t, int inverted)
{
    for (int64_t i = 0; i < n_px; ++i) {
        int d = (int)src[i] - (int)mean[i];
        dst[i] = (uint8_t)((inverted ? (d <= t) : (d > t)) ? 255 : 0);
    }
}
