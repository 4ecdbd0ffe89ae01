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
    void relocate(ysmr::LabelFrame &, uint32_t) const {}
    uint32_t exclusive_scan(uint32_t *a, int n) const
    {
        uint32_t s = 0;
        for (int i = 0; i < n; ++i) { uint32_t v = a[i]; a[i] = s; s += v; }
        return s;
    }
    uint32_t atomic_min(uint32_t *p, uint32_t v) const { uint32_t o = *p; if (v < o) *p = v; return o; }
    void atomic_add(uint32_t *p, uint32_t v) const { *p += v; }
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

#include "../../ysmr_b200/csrc/link.cuh"

struct HostLinkCta : HostCta {
    unsigned long long atomic_min_u64(unsigned long long *p, unsigned long long v) const { const unsigned long long o = *p; if (v < o) *p = v; return o; }
    void atomic_min_i32(int32_t *p, int32_t v) const { if (v < *p) *p = v; }
    uint32_t atomic_add_u32(uint32_t *p, uint32_t v) const { const uint32_t o = *p; *p = o + v; return o; }
    bool any(int v) const { return v != 0; }
    void stage_detections(const LinkConfig &c, const LinkScratch &x, const FrameScratch &f, const float *dets, int m, DetGrid &G)
    {
        stage_detections_generic(*this, c, x, f, dets, m, G);
    }
};

struct HostLinker {
    LinkConfig c;
    LinkState s;
    LinkScratch x;
    std::vector<std::vector<double>> gains;
    std::vector<int32_t> hdr, order0, order1, free_slots, id, gone, mode, hist_n, col_row, row_arg, list, table, cell_items, fflags;
    std::vector<uint32_t> cell_start;
    std::vector<float2> dxy;
    FrameScratch f;
    std::vector<double> px, py, hist, wgt, xh, row_min;
    double exp_tab[NP_EXP_TABLE];
    std::vector<float> iw, ih, ideg;
    std::vector<unsigned long long> col_best;
    std::vector<uint32_t> flag;
};

extern "C" {

// gains: for each filter 4 x (2*n_i) row-major doubles, concatenated
void *emul_link_create(double fps, int use_gsff, int n_f, const int32_t *n_i, const double *gain_full, int max_tracks,
                       int max_blobs, double max_distance)
{
    HostLinker *L = new HostLinker();
    LinkConfig &c = L->c;
    c.max_disappeared = fps; c.max_distance = max_distance; c.use_gsff = use_gsff; c.n_f = n_f;
    c.max_tracks = max_tracks; c.max_blobs = max_blobs;
    c.cross_zero = 1; c.xy_same = 1;
    c.grid_cell = 39.0; c.grid_w = 32; c.grid_h = 24;                  // what ysmr_create derives for 1228 x 922
    const double *g = gain_full;
    L->gains.resize(n_f);
    for (int i = 0; i < n_f; ++i) {
        c.n_i[i] = n_i[i];
        const int n = n_i[i];
        L->gains[i].resize(4 * n);
        for (int k = 0; k < n; ++k) {
            L->gains[i][k] = g[0 * 2 * n + 2 * k];
            L->gains[i][n + k] = g[0 * 2 * n + 2 * k + 1];
            L->gains[i][2 * n + k] = g[1 * 2 * n + 2 * k];
            L->gains[i][3 * n + k] = g[1 * 2 * n + 2 * k + 1];
            if (L->gains[i][n + k] != 0.0 || L->gains[i][2 * n + k] != 0.0) c.cross_zero = 0;
            if (L->gains[i][k] != L->gains[i][3 * n + k]) c.xy_same = 0;
        }
        c.gain[i] = L->gains[i].data();
        g += 4 * 2 * n;
    }
    c.hist_len = use_gsff ? c.n_i[n_f - 1] + 1 : 1;
    for (int k = 0; k < NP_EXP_TABLE; ++k) L->exp_tab[k] = d_from_bits(np_exp_table_bits(k));
    c.exp_tab = L->exp_tab;
    const int T = max_tracks, B = max_blobs;
    L->hdr.assign(8, 0); L->order0.resize(T); L->order1.resize(T); L->free_slots.resize(T);
    L->id.assign(T, -1); L->gone.assign(T, 0); L->mode.assign(T, 0); L->hist_n.assign(T, 0);
    L->px.resize(T); L->py.resize(T); L->iw.resize(T); L->ih.resize(T); L->ideg.resize(T);
    L->hist.resize((size_t)T * c.hist_len * 2); L->wgt.resize((size_t)T * LINK_MAX_FILTERS); L->xh.resize((size_t)T * LINK_MAX_FILTERS * 2);
    for (int i = 0; i < T; ++i) L->free_slots[i] = T - 1 - i;
    L->hdr[2] = T;
    LinkState &s = L->s;
    s.hdr = L->hdr.data(); s.order[0] = L->order0.data(); s.order[1] = L->order1.data(); s.free_slots = L->free_slots.data();
    s.id = L->id.data(); s.px = L->px.data(); s.py = L->py.data(); s.iw = L->iw.data(); s.ih = L->ih.data(); s.ideg = L->ideg.data();
    s.gone = L->gone.data(); s.mode = L->mode.data(); s.hist_n = L->hist_n.data();
    s.hist = L->hist.data(); s.wgt = L->wgt.data(); s.xh = L->xh.data();
    L->col_best.resize(B); L->col_row.resize(B); L->row_min.resize(T); L->row_arg.resize(T);
    L->flag.resize((T > B ? T : B) + 2); L->list.resize(B); L->table.resize(set_table_capacity(B));
    LinkScratch &x = L->x;
    x.row_min = L->row_min.data(); x.row_arg = L->row_arg.data();
    L->dxy.resize(B); L->cell_items.resize(B); L->cell_start.resize(LINK_GRID_CELLS + 2); L->fflags.assign(4, 0);
    L->f.dxy = L->dxy.data(); L->f.col_best = L->col_best.data(); L->f.col_row = L->col_row.data();
    L->f.cell_items = L->cell_items.data(); L->f.cell_start = L->cell_start.data(); L->f.flags = L->fflags.data();
    x.flag = L->flag.data(); x.list = L->list.data(); x.table = L->table.data(); x.set_table_size = (int)L->table.size();
    return L;
}

void emul_link_destroy(void *h) { delete (HostLinker *)h; }

void emul_link_set_grid(void *h, double cell, int gw, int gh)
{
    HostLinker *L = (HostLinker *)h;
    L->c.grid_cell = cell; L->c.grid_w = gw; L->c.grid_h = gh;
}

// blob_count[n_frames], blobs[n_frames][max_blobs][5]; rows as 5 doubles? -> RowOut array.  Returns status bits.
int emul_link_chunk(void *h, const int32_t *blob_count, const float *blobs, int first_frame, int n_frames, void *rows,
                    long long rows_capacity, long long *n_rows)
{
    HostLinker *L = (HostLinker *)h;
    int32_t status = 0, first_bad = 0x7fffffff;
    LinkIo io;
    io.blob_count = blob_count; io.blobs = blobs; io.rows = (RowOut *)rows; io.rows_capacity = rows_capacity;
    io.n_rows = n_rows; io.append = 0; io.status = &status; io.first_bad = &first_bad;
    HostLinkCta cta;
    link_chunk(cta, L->c, L->s, L->x, L->f, io, first_frame, n_frames);
    return status;
}

// numpy.exp restatement (link.cuh: np_exp_nonpos) on an array of non-positive arguments
void emul_np_exp(const double *x, double *out, int n)
{
    double tab[NP_EXP_TABLE];
    for (int k = 0; k < NP_EXP_TABLE; ++k) tab[k] = d_from_bits(np_exp_table_bits(k));
    for (int i = 0; i < n; ++i) out[i] = np_exp_nonpos(x[i], tab);
}

// one row of numpy.dot(gain (rows x 2n), y (2n)) in the dgemv_t order (link.cuh: blas_row_dot); g_row = 2n doubles
double emul_blas_row_dot(const double *g_row, const double *y, int n)
{
    std::vector<double> g0(n), g1(n);
    for (int k = 0; k < n; ++k) { g0[k] = g_row[2 * k]; g1[k] = g_row[2 * k + 1]; }
    return blas_row_dot(g0.data(), g1.data(), false, y, 2, 0, n, n);
}

void emul_set_order(int32_t *keys, int n)
{
    std::vector<int32_t> table(set_table_capacity(n));
    cpython_set_order(keys, n, table.data());
}

}  // extern "C"

#include "../../ysmr_b200/csrc/select.cuh"

extern "C" {

// ndarray.sum() / n as csrc/select.cuh restates it (checked against numpy / pandas in tests/test_select_host.py)
double emul_np_mean(const double *a, int64_t n) { return np_mean(a, n); }
double emul_np_sum(const double *a, int64_t n) { return np_pairwise_sum(a, n); }

// find_good_tracks for every track of a cleaned-up frame (single lane).  cfg: the SelectCfg fields in declaration order
// as doubles.  Outputs per track: good_start, good_stop (-1 = none), kick reason.
void emul_select_tracks(const int32_t *track_start, int n_tracks, int n_rows, const uint32_t *t, const double *x, const double *y,
                        const double *area, const double *ratio, const int8_t *outlier, const double *cfg, int32_t *good_start,
                        int32_t *good_stop, int32_t *kick)
{
    SelectCols c{t, x, y, area, ratio, outlier};
    SelectCfg g{};
    g.min_len = (int)cfg[0]; g.max_holes = (int)cfg[1]; g.max_recursion = (int)cfg[2]; g.max_empty = cfg[3]; g.lower = cfg[4];
    g.upper = cfg[5]; g.ratio_min = cfg[6]; g.ratio_max = cfg[7]; g.edge = cfg[8]; g.frame_h = (int)cfg[9]; g.frame_w = (int)cfg[10];
    g.limit_frames = (int)cfg[11]; g.limit_exactly = (int)cfg[12];
    std::vector<SelectSeg> stack(g.max_recursion + 2);
    const OneLane one;
    for (int k = 0; k < n_tracks; ++k) {
        const int lo = track_start[k], hi = (k + 1 < n_tracks ? track_start[k + 1] : n_rows) - 1;
        int gs = -1, ge = -1;
        kick[k] = select_track(one, c, g, lo, hi, stack.data(), (int)stack.size(), &gs, &ge);
        good_start[k] = gs; good_stop[k] = ge;
    }
}

}  // extern "C"

#include "../../ysmr_b200/csrc/stats.cuh"

extern "C" {

// evaluate_tracks' per-track reductions (csrc/stats.cuh) for every track of a selected frame, single thread
void emul_track_statistics(const int32_t *track_start, int n_tracks, int n_rows, const uint32_t *t, const double *x, const double *y,
                           const double *w, const double *h, double px, double fps, int kernel2, double *out)
{
    StatCols c{t, x, y, w, h};
    StatCfg g{px, fps, kernel2};
    std::vector<uint8_t> ma(n_rows), mb(n_rows);
    for (int k = 0; k < n_tracks; ++k) {
        const int lo = track_start[k], hi = (k + 1 < n_tracks ? track_start[k + 1] : n_rows) - 1;
        track_statistics_serial(c, g, lo, hi, ma.data(), mb.data(), out + (size_t)k * STAT_COLUMNS);
    }
}

}  // extern "C"
