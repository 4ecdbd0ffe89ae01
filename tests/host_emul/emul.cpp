// TEST INFRASTRUCTURE ONLY.  Compiles the YSMR_HD logic of ysmr_b200/csrc/*.cuh for the host CPU so that the
// CPU-only test suite can check it against cv2/scipy and the oracle without a GPU.  Never loaded by ysmr_b200/.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../ysmr_b200/csrc/common.cuh"
#include "../../ysmr_b200/csrc/geometry.cuh"

using namespace ysmr;

static std::vector<uint32_t> pack_bits(const uint8_t *img, int h, int w)
{
    int ww = words_per_row(w);
    std::vector<uint32_t> bits((size_t)h * ww, 0u);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            if (img[(size_t)y * w + x]) bits[(size_t)y * ww + (x >> 5)] |= 1u << (x & 31);
    return bits;
}

extern "C" {

// contour vertices of the blob starting at (x0,y0); returns count (writes at most cap points as x,y int32 pairs)
int emul_trace(const uint8_t *img, int h, int w, int x0, int y0, int32_t *xy, int cap)
{
    auto bits = pack_bits(img, h, w);
    BitImage bi{bits.data(), h, w, words_per_row(w)};
    std::vector<Pt16> pts(cap > 0 ? cap : 1);
    int n = trace_outer_border(bi, x0, y0, pts.data(), cap);
    for (int i = 0; i < n && i < cap; ++i) { xy[2 * i] = pts[i].x; xy[2 * i + 1] = pts[i].y; }
    return n;
}

// hull (as points) of given contour points
int emul_hull(const int32_t *xy, int n, int32_t *hxy)
{
    std::vector<Pt16> pts(n);
    for (int i = 0; i < n; ++i) { pts[i].x = (int16_t)xy[2 * i]; pts[i].y = (int16_t)xy[2 * i + 1]; }
    std::vector<uint16_t> ord(n + 4), st(n + 4), hull(n + 4);
    int nh = convex_hull(pts.data(), n, ord.data(), st.data(), hull.data());
    for (int i = 0; i < nh; ++i) { hxy[2 * i] = pts[hull[i]].x; hxy[2 * i + 1] = pts[hull[i]].y; }
    return nh;
}

// min-area rects for n blobs given their raster-first pixels
void emul_blob_rects(const uint8_t *img, int h, int w, const int32_t *first_xy, int n, float *out, int32_t *npts)
{
    auto bits = pack_bits(img, h, w);
    BitImage bi{bits.data(), h, w, words_per_row(w)};
    for (int i = 0; i < n; ++i) {
        int cap = 64;
        for (;;) {
            std::vector<Pt16> pts(cap);
            std::vector<uint16_t> ord(cap + 4), st(cap + 4), hull(cap + 4);
            int c = blob_rect(bi, first_xy[2 * i], first_xy[2 * i + 1], pts.data(), ord.data(), st.data(), hull.data(), cap,
                              out + 5 * i);
            if (c <= cap) { npts[i] = c; break; }
            cap = c;
        }
    }
}

}  // extern "C"

#include "../../ysmr_b200/csrc/label.cuh"

struct HostCta {
    int tid() const { return 0; }
    int nthr() const { return 1; }
    void sync() const {}
    uint32_t exclusive_scan(uint32_t *a, int n) const
    {
        uint32_t s = 0;
        for (int i = 0; i < n; ++i) { uint32_t v = a[i]; a[i] = s; s += v; }
        return s;
    }
    uint32_t atomic_min(uint32_t *p, uint32_t v) const { uint32_t o = *p; if (v < o) *p = v; return o; }
    void atomic_and(uint32_t *p, uint32_t m) const { *p &= m; }
    void atomic_or_i32(int32_t *p, int32_t v) const { *p |= v; }
};

extern "C" {

// mask/markers: h*w bytes (non-zero = set); markers may be NULL (DIRECT mode).  out: h*w bytes {0,255}.
// first_xy: max_blobs x 2 int32.  Returns status bits.
int emul_label(const uint8_t *mask, const uint8_t *markers, int h, int w, int max_runs, int max_blobs, uint8_t *out,
               int32_t *first_xy, int32_t *count, uint32_t *counts)
{
    int ww = words_per_row(w);
    auto img = pack_bits(mask, h, w);
    std::vector<uint32_t> seed;
    if (markers) seed = pack_bits(markers, h, w);
    std::vector<uint32_t> row_start(h + 1), krow_start(h + 1), parent(max_runs), sd(max_runs), kparent(max_runs),
        gparent(max_runs + h + 2), ext(max_runs), fxy(max_blobs);
    std::vector<uint16_t> rx0(max_runs), rx1(max_runs), ry(max_runs), kx0(max_runs), kx1(max_runs), ky(max_runs);
    int32_t status = 0;
    LabelFrame f;
    f.h = h; f.w = w; f.ww = ww; f.max_runs = max_runs; f.max_blobs = max_blobs;
    f.mode_propagate = markers ? 1 : 0;
    f.img = img.data(); f.seedimg = markers ? seed.data() : nullptr;
    f.row_start = row_start.data();
    f.rx0 = rx0.data(); f.rx1 = rx1.data(); f.ry = ry.data();
    f.parent = parent.data(); f.seed = sd.data();
    if (markers) { f.krow_start = krow_start.data(); f.kx0 = kx0.data(); f.kx1 = kx1.data(); f.ky = ky.data(); }
    else { f.krow_start = row_start.data(); f.kx0 = rx0.data(); f.kx1 = rx1.data(); f.ky = ry.data(); }
    f.kparent = kparent.data(); f.gparent = gparent.data(); f.ext = ext.data();
    f.blob_count = count; f.first_xy = fxy.data(); f.status = &status; f.counts = counts;
    HostCta cta;
    label_frame(cta, f);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) out[(size_t)y * w + x] = (img[(size_t)y * ww + (x >> 5)] >> (x & 31)) & 1u ? 255 : 0;
    for (int i = 0; i < *count; ++i) { first_xy[2 * i] = fxy[i] & 0xFFFF; first_xy[2 * i + 1] = fxy[i] >> 16; }
    return status;
}

}  // extern "C"
