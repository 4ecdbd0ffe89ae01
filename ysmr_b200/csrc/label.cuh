// K2 -- run-based union-find labelling of one frame's packed threshold masks.
//
// Replaces, per frame, scipy.ndimage.binary_propagation(markers, mask=thresh) (/root/reference/ysmr/
// track_eval.py:211-214) and the component discovery part of cv2.findContours(thresh, RETR_EXTERNAL, ...)
// (track_eval.py:273): which 8-connected components exist, which of them are external (not nested in a hole of
// another component), the raster-first pixel of each, and cv2's output order (descending raster order of that
// pixel).  The border following itself is K3 (geometry.cuh).
//
// Representation: a frame is a list of horizontal runs (maximal segments of 1s) in raster order.  Because the run
// index is the raster order, "smallest index in a component" is the run that starts at the component's raster-first
// pixel, so union-by-minimum-index gives first pixels and cv2's ordering for free.
//
//   mode PROPAGATE (adaptive double threshold, markers subset of mask):
//     runs of `mask` -> 4-connected union-find -> a component is kept iff one of its runs covers a marker pixel
//      -> the bits of the other components are cleared in place (mask becomes the image handed to
//     findContours) -> kept runs are compacted.
//   mode DIRECT (single threshold, mean/std mode, or the dark-on-light quirk where mask is a subset of markers and
//     binary_propagation returns the marker image, SURVEY finding 8): the runs of the one image are the kept runs.
//   then: 8-connected union-find over kept runs; 4-connected union-find over background gaps with a virtual OUTER
//   node for everything that touches the zero padding; a component is external iff the gap left of its first
//   pixel is OUTER; external roots are emitted in descending index order.
//
// Every phase is a loop "for (i = cta.tid(); i < n; i += cta.nthr())" separated by cta.sync(); the Cta policy
// supplies the barrier, a block-wide exclusive scan and atomicMin.  The device policy is in label.cu; tests/host_emul
// provides a sequential one, so the logic below is exercised on the CPU against scipy/cv2.
#pragma once
#include "common.cuh"

namespace ysmr {

struct LabelFrame {
    // geometry
    int h, w, ww;
    int max_runs, max_blobs;
    int mode_propagate;          // 1 = PROPAGATE, 0 = DIRECT
    // images (packed). `img` is runs' source and becomes the findContours image; `seedimg` only for PROPAGATE.
    uint32_t *img;
    const uint32_t *seedimg;
    // scratch, per frame
    uint32_t *row_start;         // [h + 1]   first run index of each row (source runs)
    uint32_t *krow_start;        // [h + 1]   same for kept runs (aliases row_start in DIRECT mode)
    uint16_t *rx0, *rx1, *ry;    // [max_runs] source runs
    uint32_t *parent;            // [max_runs] 4-conn union-find, later kept flag / kept prefix
    uint32_t *seed;              // [max_runs]
    uint16_t *kx0, *kx1, *ky;    // [max_runs] kept runs (alias the source arrays in DIRECT mode)
    uint32_t *kparent;           // [max_runs] 8-conn union-find
    uint32_t *gparent;           // [max_runs + h + 2] gap union-find, node 0 = OUTER
    uint32_t *ext;               // [max_runs] external-root flag, later its prefix
    // outputs
    int32_t *blob_count;         // [1]
    uint32_t *first_xy;          // [max_blobs] x | y << 16, cv2 order
    int32_t *status;             // [1] YSMR_ST_* bits (atomicOr)
    uint32_t *counts;            // [4] n_runs, n_kept, n_ext (for tests / profiling)
};

enum { LABEL_ST_RUN_OVERFLOW = 1, LABEL_ST_BLOB_OVERFLOW = 2 };

// ---- run extraction ---------------------------------------------------------------------------------------------

YSMR_HD int ctz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

YSMR_HD int popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// runs of one word appended at index k (continuing a run that is open from the previous word)
YSMR_HD void write_word_runs(uint32_t b, int base, int y, uint32_t &k, bool &open, uint32_t cap, uint16_t *x0s, uint16_t *x1s,
                             uint16_t *ys)
{
    if (open) {
        if (b == 0xFFFFFFFFu) return;
        const int z = ctz32(~b);                           // first zero bit closes the run
        if (k < cap) x1s[k] = (uint16_t)(base + z - 1);
        ++k;
        open = false;
        b &= ~((1u << z) - 1u);
    }
    while (b) {
        const int s = ctz32(b);
        const uint32_t t = ~(b >> s);                      // zeros of the shifted word; bits above 31-s are 1
        if (k < cap) { x0s[k] = (uint16_t)(base + s); ys[k] = (uint16_t)y; }
        const int len = t ? ctz32(t) : 32;
        if (s + len >= 32) { open = true; break; }         // runs to the end of this word
        if (k < cap) x1s[k] = (uint16_t)(base + s + len - 1);
        ++k;
        b &= ~(((1u << len) - 1u) << s);
    }
}

// write the runs of one row at index k.., returns the next free index.  Writes nothing at or beyond cap.
YSMR_HD uint32_t write_row_runs(const uint32_t *row, int ww, int w, int y, uint32_t k, uint32_t cap, uint16_t *x0s,
                                uint16_t *x1s, uint16_t *ys)
{
    bool open = false;
    int i = 0;
    for (; i + 4 <= ww; i += 4) {
        const uint32_t b0 = row[i], b1 = row[i + 1], b2 = row[i + 2], b3 = row[i + 3];
        if (!open && (b0 | b1 | b2 | b3) == 0u) continue;
        write_word_runs(b0, i << 5, y, k, open, cap, x0s, x1s, ys);
        write_word_runs(b1, (i + 1) << 5, y, k, open, cap, x0s, x1s, ys);
        write_word_runs(b2, (i + 2) << 5, y, k, open, cap, x0s, x1s, ys);
        write_word_runs(b3, (i + 3) << 5, y, k, open, cap, x0s, x1s, ys);
    }
    for (; i < ww; ++i) write_word_runs(row[i], i << 5, y, k, open, cap, x0s, x1s, ys);
    if (open) {
        if (k < cap) x1s[k] = (uint16_t)(w - 1);
        ++k;
    }
    return k;
}

// ---- union-find (union by minimum index; lock-free with atomicMin on the device) ----------------------------------

template <class Cta>
YSMR_HD uint32_t uf_find(const Cta &cta, volatile uint32_t *parent, uint32_t i)
{
    uint32_t p = parent[i];
    while (p != i) { i = p; p = parent[i]; }
    return i;
}

template <class Cta>
YSMR_HD void uf_union(const Cta &cta, uint32_t *parent, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(cta, parent, a);
        b = uf_find(cta, parent, b);
        if (a == b) return;
        if (a > b) { uint32_t t = a; a = b; b = t; }       // a < b: hang b under a
        const uint32_t old = cta.atomic_min(&parent[b], a);
        if (old == b) return;
        b = old;                                           // somebody else re-parented b meanwhile: retry
    }
}

// first run j in [lo, hi) with x1s[j] >= x  (runs of a row are sorted and disjoint)
YSMR_HD uint32_t lower_run(const uint16_t *x1s, uint32_t lo, uint32_t hi, int x)
{
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((int)x1s[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// link run i with the overlapping runs of the previous row; reach = 0 for 4-connectivity, 1 for 8-connectivity
template <class Cta>
YSMR_HD void link_up(const Cta &cta, uint32_t i, const uint16_t *x0s, const uint16_t *x1s, const uint16_t *ys,
                     const uint32_t *row_start, uint32_t *parent, int reach)
{
    const int y = ys[i];
    if (y == 0) return;
    const uint32_t lo = row_start[y - 1], hi = row_start[y];
    if (lo == hi) return;
    const int a0 = (int)x0s[i] - reach, a1 = (int)x1s[i] + reach;
    for (uint32_t j = lower_run(x1s, lo, hi, a0); j < hi && (int)x0s[j] <= a1; ++j) uf_union(cta, parent, i, j);
}

YSMR_HD bool any_bits(const uint32_t *row, int x0, int x1)
{
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    for (int i = w0; i <= w1; ++i) {
        uint32_t m = 0xFFFFFFFFu;
        if (i == w0) m &= 0xFFFFFFFFu << (x0 & 31);
        if (i == w1) m &= 0xFFFFFFFFu >> (31 - (x1 & 31));
        if (row[i] & m) return true;
    }
    return false;
}

template <class Cta>
YSMR_HD void clear_bits(const Cta &cta, uint32_t *row, int x0, int x1)
{
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    for (int i = w0; i <= w1; ++i) {
        uint32_t m = 0xFFFFFFFFu;
        if (i == w0) m &= 0xFFFFFFFFu << (x0 & 31);
        if (i == w1) m &= 0xFFFFFFFFu >> (31 - (x1 & 31));
        cta.atomic_and(&row[i], ~m);
    }
}

// ---- the per-frame program ---------------------------------------------------------------------------------------

template <class Cta>
YSMR_HD void label_frame(Cta &cta, const LabelFrame &f0)
{
    const int tid = cta.tid(), nthr = cta.nthr();
    LabelFrame f = f0;
    const uint32_t cap = (uint32_t)f.max_runs;

    // P1: runs per row.  The CTA streams the bit image once, word by word (consecutive threads read consecutive words); the
    // masks are sparse, so only the few non-zero words look at their left neighbour and add to their row's counter.
    for (int y = tid; y <= f.h; y += nthr) f.row_start[y] = 0;
    cta.sync();
    {
        const int n_words = f.h * f.ww;
        for (int base = tid; base < n_words; base += 4 * nthr) {     // four independent loads in flight per thread
            uint32_t b[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int k = 0; k < 4; ++k) { const int i = base + k * nthr; b[k] = i < n_words ? f.img[i] : 0u; }
            if ((b[0] | b[1] | b[2] | b[3]) == 0u) continue;
            for (int k = 0; k < 4; ++k) {
                if (b[k] == 0u) continue;
                const int i = base + k * nthr;
                const int y = i / f.ww, xw = i - y * f.ww;
                const uint32_t carry = xw ? f.img[i - 1] >> 31 : 0u;
                const uint32_t n = (uint32_t)popc32(b[k] & ~((b[k] << 1) | carry));
                if (n) cta.atomic_add(&f.row_start[y], n);
            }
        }
    }
    cta.sync();
    uint32_t n_runs = cta.exclusive_scan(f.row_start, f.h + 1);
    bool overflow = n_runs > cap;
    if (overflow) {
        if (tid == 0) { cta.atomic_or_i32(f.status, LABEL_ST_RUN_OVERFLOW); *f.blob_count = 0; f.counts[0] = n_runs; }
        return;
    }
    // The union-find arrays move to CTA-local fast memory when the frame's runs fit there (device policy: shared memory).
    cta.relocate(f, n_runs);
    // P2: write runs
    for (int y = tid; y < f.h; y += nthr)
        if (f.row_start[y + 1] != f.row_start[y])                   // most rows are empty
            write_row_runs(f.img + (int64_t)y * f.ww, f.ww, f.w, y, f.row_start[y], cap, f.rx0, f.rx1, f.ry);
    cta.sync();

    uint32_t nk = n_runs;
    if (f.mode_propagate) {
        // P3: 4-connected components of the mask
        for (uint32_t i = tid; i < n_runs; i += nthr) { f.parent[i] = i; f.seed[i] = 0; }
        cta.sync();
        for (uint32_t i = tid; i < n_runs; i += nthr) link_up(cta, i, f.rx0, f.rx1, f.ry, f.row_start, f.parent, 0);
        cta.sync();
        // P4: a component is seeded when any of its runs covers a marker pixel
        for (uint32_t i = tid; i < n_runs; i += nthr)
            if (any_bits(f.seedimg + (int64_t)f.ry[i] * f.ww, f.rx0[i], f.rx1[i])) f.seed[uf_find(cta, f.parent, i)] = 1;
        cta.sync();
        // P5: kept flag per run (stored over kparent as 0/1), unkept runs vanish from the image
        for (uint32_t i = tid; i < n_runs; i += nthr) {
            const uint32_t keep = f.seed[uf_find(cta, f.parent, i)];
            f.kparent[i] = keep;
            if (!keep) clear_bits(cta, f.img + (int64_t)f.ry[i] * f.ww, f.rx0[i], f.rx1[i]);
        }
        cta.sync();
        nk = cta.exclusive_scan(f.kparent, (int)n_runs);          // kparent[i] = kept runs before i
        // P6: compact kept runs, rebuild the row index
        for (uint32_t i = tid; i < n_runs; i += nthr) {
            const uint32_t r = uf_find(cta, f.parent, i);
            if (f.seed[r]) {
                const uint32_t k = f.kparent[i];
                f.kx0[k] = f.rx0[i]; f.kx1[k] = f.rx1[i]; f.ky[k] = f.ry[i];
            }
        }
        for (int y = tid; y <= f.h; y += nthr) {
            const uint32_t s = f.row_start[y];
            f.krow_start[y] = s < n_runs ? f.kparent[s] : nk;
        }
        cta.sync();
    }
    // (in DIRECT mode kx0/kx1/ky/krow_start alias the source arrays)

    // P7: 8-connected components of the kept runs
    for (uint32_t i = tid; i < nk; i += nthr) f.kparent[i] = i;
    // gaps: node 0 = OUTER; gap k of row y (left of the k-th kept run of that row, or right of the last) = ks+k+y+1
    const uint32_t n_gap = nk + (uint32_t)f.h + 1;
    for (uint32_t g = tid; g < n_gap; g += nthr) f.gparent[g] = g;
    cta.sync();
    for (uint32_t i = tid; i < nk; i += nthr) link_up(cta, i, f.kx0, f.kx1, f.ky, f.krow_start, f.kparent, 1);
    // P8: background gaps.  One work item per (row, k): rows enumerate their gaps themselves.
    for (int y = tid; y < f.h; y += nthr) {
        const uint32_t ks = f.krow_start[y], ke = f.krow_start[y + 1];
        const uint32_t cnt = ke - ks;
        for (uint32_t k = 0; k <= cnt; ++k) {
            const int a0 = k == 0 ? 0 : (int)f.kx1[ks + k - 1] + 1;
            const int a1 = k == cnt ? f.w - 1 : (int)f.kx0[ks + k] - 1;
            if (a0 > a1) continue;                                  // empty gap (run touches the image edge)
            const uint32_t g = ks + k + (uint32_t)y + 1;
            if (y == 0 || y == f.h - 1 || k == 0 || k == cnt) {     // touches the zero padding
                uf_union(cta, f.gparent, g, 0u);
                if (y == 0) continue;
            }
            // overlapping gaps of the previous row (4-connectivity: share a column)
            const uint32_t ps = f.krow_start[y - 1], pe = ks, pcnt = pe - ps;
            // first gap q of the previous row whose end >= a0: first run with x0 > a0 (gap q ends at x0[q]-1)
            uint32_t lo = 0, hi = pcnt;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if ((int)f.kx0[ps + mid] - 1 < a0) lo = mid + 1; else hi = mid;
            }
            for (uint32_t q = lo; q <= pcnt; ++q) {
                const int b0 = q == 0 ? 0 : (int)f.kx1[ps + q - 1] + 1;
                if (b0 > a1) break;
                const int b1 = q == pcnt ? f.w - 1 : (int)f.kx0[ps + q] - 1;
                if (b0 > b1) continue;
                uf_union(cta, f.gparent, g, ps + q + (uint32_t)(y - 1) + 1);
            }
        }
    }
    cta.sync();
    // P9: external roots
    for (uint32_t i = tid; i < nk; i += nthr) {
        uint32_t e = 0;
        if (uf_find(cta, f.kparent, i) == i) {
            const int y = f.ky[i];
            if (f.kx0[i] == 0) e = 1;                               // left neighbour is the padding itself
            else e = uf_find(cta, f.gparent, i + (uint32_t)y + 1) == 0u;
        }
        f.ext[i] = e;
    }
    cta.sync();
    const uint32_t n_ext = cta.exclusive_scan(f.ext, (int)nk);       // ext[i] = external roots before i
    // P10: emit first pixels, last root first (cv2.findContours order)
    const uint32_t n_out = n_ext > (uint32_t)f.max_blobs ? (uint32_t)f.max_blobs : n_ext;
    for (uint32_t i = tid; i < nk; i += nthr) {
        const uint32_t before = f.ext[i];
        const uint32_t after = (i + 1 < nk) ? f.ext[i + 1] : n_ext;
        if (after != before) {
            const uint32_t pos = n_ext - 1 - before;
            if (pos < n_out) f.first_xy[pos] = (uint32_t)f.kx0[i] | ((uint32_t)f.ky[i] << 16);
        }
    }
    if (tid == 0) {
        *f.blob_count = (int32_t)n_out;
        if (n_ext > (uint32_t)f.max_blobs) cta.atomic_or_i32(f.status, LABEL_ST_BLOB_OVERFLOW);
        f.counts[0] = n_runs; f.counts[1] = nk; f.counts[2] = n_ext;
    }
}

}  // namespace ysmr
