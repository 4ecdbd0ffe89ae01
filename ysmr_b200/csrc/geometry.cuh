// K3 -- per-blob geometry: outer border trace -> convex hull -> rotating calipers -> min-area rectangle.
//
// Replaces, for one blob, the pair of calls at /root/reference/ysmr/track_eval.py:273 (cv2.findContours with
// RETR_EXTERNAL, CHAIN_APPROX_SIMPLE -- here only the border following of ONE outer border; which blobs are
// external and their order comes from the labelling kernel) and track_eval.py:287 (cv2.minAreaRect).  The
// arithmetic follows OpenCV 4.13's published algorithms (Suzuki-Abe border following; Sklansky hull on points sorted
// by (x, y, index); Toussaint's rotating calipers in float32 with double only where OpenCV uses double) because the
// 1e-3 px / 1e-3 rad parity bar -- and the exact-tie behaviour -- depends on evaluation order.
//
// All functions are YSMR_HD: the device runs one thread per blob; tests/host_emul compiles the same code for the CPU.
#pragma once
#include "common.cuh"

#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

namespace ysmr {

struct Pt16 {
    int16_t x, y;
};

// 8-neighbourhood, counter-clockwise on screen starting east; y grows downwards.
YSMR_HD int dir_dx(int d) { return (d == 0 || d == 1 || d == 7) ? 1 : ((d >= 3 && d <= 5) ? -1 : 0); }
YSMR_HD int dir_dy(int d) { return (d >= 1 && d <= 3) ? -1 : ((d >= 5 && d <= 7) ? 1 : 0); }

// Follow the outer border of the 8-connected component whose raster-first pixel is (x0, y0).
// Emits the CHAIN_APPROX_SIMPLE vertices (a point is kept when the chain direction changes) into pts[0..cap);
// returns the number of vertices the contour has (which may exceed cap: the caller must then retry with more room).
YSMR_HD int trace_outer_border(const BitImage &img, int x0, int y0, Pt16 *pts, int cap)
{
    // first neighbour clockwise from west: NW, N, NE, E, SE, S, SW
    int s = 4;
    bool found = false;
    for (int k = 0; k < 7; ++k) {
        s = (s - 1) & 7;
        if (img.at(x0 + dir_dx(s), y0 + dir_dy(s))) { found = true; break; }
    }
    if (!found) {
        if (cap > 0) { pts[0].x = (int16_t)x0; pts[0].y = (int16_t)y0; }
        return 1;
    }
    const int fx = x0 + dir_dx(s), fy = y0 + dir_dy(s);   // the pixel we must be leaving from when we close
    int cx = x0, cy = y0;
    int prev = s ^ 4;
    int n = 0;
    for (;;) {
        // next border pixel: first foreground neighbour counter-clockwise after direction s
        int d = s, nx = cx, ny = cy;
        for (int k = 0; k < 8; ++k) {
            d = (d + 1) & 7;
            nx = cx + dir_dx(d); ny = cy + dir_dy(d);
            if (img.at(nx, ny)) break;
        }
        if (d != prev) {
            if (n < cap) { pts[n].x = (int16_t)cx; pts[n].y = (int16_t)cy; }
            ++n;
            prev = d;
        }
        if (nx == x0 && ny == y0 && cx == fx && cy == fy) return n;
        cx = nx; cy = ny;
        s = (d + 4) & 7;
    }
}

// ---- convex hull, OpenCV convexHull(points, clockwise=false) order ---------------------------------------------

YSMR_HD bool hull_less(const Pt16 *p, int a, int b)
{
    if (p[a].x != p[b].x) return p[a].x < p[b].x;
    if (p[a].y != p[b].y) return p[a].y < p[b].y;
    return a < b;
}

YSMR_HD void hull_sort(const Pt16 *p, uint16_t *ord, int n)
{
    if (n <= 24) {
        for (int i = 1; i < n; ++i) {
            uint16_t v = ord[i];
            int j = i - 1;
            while (j >= 0 && hull_less(p, v, ord[j])) { ord[j + 1] = ord[j]; --j; }
            ord[j + 1] = v;
        }
        return;
    }
    // heapsort (total order, so stability is irrelevant)
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && hull_less(p, ord[child], ord[child + 1])) ++child;
            if (!hull_less(p, ord[root], ord[child])) break;
            uint16_t t = ord[root]; ord[root] = ord[child]; ord[child] = t;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; --end) {
        uint16_t t = ord[0]; ord[0] = ord[end]; ord[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && hull_less(p, ord[child], ord[child + 1])) ++child;
            if (!hull_less(p, ord[root], ord[child])) break;
            uint16_t t2 = ord[root]; ord[root] = ord[child]; ord[child] = t2;
            root = child;
        }
    }
}

YSMR_HD int isign(int v) { return (v > 0) - (v < 0); }

// Sklansky scan over the sorted points from `start` towards `end`; stack receives positions in the sorted order.
YSMR_HD int sklansky(const Pt16 *p, const uint16_t *ord, int start, int end, uint16_t *stack, int nsign, int sign2)
{
    const int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int size = 3;
    if (start == end || (p[ord[start]].x == p[ord[end]].x && p[ord[start]].y == p[ord[end]].y)) {
        stack[0] = (uint16_t)start;
        return 1;
    }
    stack[0] = (uint16_t)pprev; stack[1] = (uint16_t)pcur; stack[2] = (uint16_t)pnext;
    end += incr;
    while (pnext != end) {
        const int cury = p[ord[pcur]].y;
        const int nexty = p[ord[pnext]].y;
        const int by = nexty - cury;
        if (isign(by) != nsign) {
            const int ax = p[ord[pcur]].x - p[ord[pprev]].x;
            const int bx = p[ord[pnext]].x - p[ord[pcur]].x;
            const int ay = cury - p[ord[pprev]].y;
            const int convexity = ay * bx - ax * by;
            if (isign(convexity) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur; pcur = pnext; pnext += incr;
                stack[size++] = (uint16_t)pnext;
            } else if (pprev == start) {
                pcur = pnext; stack[1] = (uint16_t)pcur;
                pnext += incr; stack[2] = (uint16_t)pnext;
            } else {
                stack[size - 2] = (uint16_t)pnext;
                pcur = pprev;
                pprev = stack[size - 4];
                --size;
            }
        } else {
            pnext += incr;
            stack[size - 1] = (uint16_t)pnext;
        }
    }
    return size - 1;
}

// Hull of pts[0..n) as indices into pts, in OpenCV's output order.  ord, stack, hull: scratch/out arrays of at
// least n + 2 entries each.  Returns the number of hull vertices.
YSMR_HD int convex_hull(const Pt16 *p, int n, uint16_t *ord, uint16_t *stack, uint16_t *hull)
{
    for (int i = 0; i < n; ++i) ord[i] = (uint16_t)i;
    hull_sort(p, ord, n);
    int lo = 0, hi = 0;
    for (int i = 1; i < n; ++i) {
        const int y = p[ord[i]].y;
        if (p[ord[lo]].y > y) lo = i;
        if (p[ord[hi]].y < y) hi = i;
    }
    if (p[ord[0]].x == p[ord[n - 1]].x && p[ord[0]].y == p[ord[n - 1]].y) {
        hull[0] = ord[0];
        return 1;
    }
    int nout = 0;
    // upper half (towards max y); with clockwise == false the two stacks swap roles
    uint16_t *s_a = stack;
    int c_a = sklansky(p, ord, 0, hi, s_a, -1, 1);
    uint16_t *s_b = stack + c_a;
    int c_b = sklansky(p, ord, n - 1, hi, s_b, -1, -1);
    {
        uint16_t *tl = s_b; int tlc = c_b;      // swapped
        uint16_t *tr = s_a; int trc = c_a;
        for (int i = 0; i < tlc - 1; ++i) hull[nout++] = ord[tl[i]];
        for (int i = trc - 1; i > 0; --i) hull[nout++] = ord[tr[i]];
        int stop = trc > 2 ? tr[1] : (tlc > 2 ? tl[tlc - 2] : -1);
        // the stacks are reused by the lower half, so keep what the collinearity test needs
        const int stop_idx = stop;
        uint16_t *bl = stack;
        int blc = sklansky(p, ord, 0, lo, bl, 1, -1);
        uint16_t *br = stack + blc;
        int brc = sklansky(p, ord, n - 1, lo, br, 1, 1);
        if (stop_idx >= 0) {
            int chk = blc > 2 ? bl[1] : (blc + brc > 2 ? br[2 - blc] : -1);
            if (chk == stop_idx || (chk >= 0 && p[ord[chk]].x == p[ord[stop_idx]].x && p[ord[chk]].y == p[ord[stop_idx]].y)) {
                blc = blc < 2 ? blc : 2;
                brc = brc < 2 ? brc : 2;
            }
        }
        for (int i = 0; i < blc - 1; ++i) hull[nout++] = ord[bl[i]];
        for (int i = brc - 1; i > 0; --i) hull[nout++] = ord[br[i]];
    }
    // cyclic shift so that the original indices run monotonically
    if (nout >= 3) {
        int mn = 0, mx = 0, lt = 0;
        for (int i = 1; i < nout; ++i) {
            const int idx = hull[i];
            lt += hull[i - 1] < idx;
            if (lt > 1 && lt <= i - 2) break;
            if (idx < hull[mn]) mn = i;
            if (idx > hull[mx]) mx = i;
        }
        int mm = mx - mn; if (mm < 0) mm = -mm;
        if ((mm == 1 || mm == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
            const bool ascending = (mx + 1) % nout == mn;
            const int i0 = ascending ? mn : mx;
            if (i0 > 0) {
                int j = i0, i = 0;
                for (; i < nout; ++i) {
                    const int cur = stack[i] = hull[j];
                    const int nj = j + 1 < nout ? j + 1 : 0;
                    const int nxt = hull[nj];
                    if (i < nout - 1 && (ascending != (cur < nxt))) break;
                    j = nj;
                }
                if (i == nout)
                    for (int k = 0; k < nout; ++k) hull[k] = stack[k];
            }
        }
    }
    return nout;
}

// ---- rotating calipers (float32, no contraction) + minAreaRect ---------------------------------------------------

struct EdgeVec {
    float vx, vy, inv;
};

YSMR_HD EdgeVec hull_edge(const Pt16 *p, const uint16_t *hull, int n, int i)
{
    const int j = (i + 1 < n) ? i + 1 : 0;
    const double dx = (double)(float)p[hull[j]].x - (double)(float)p[hull[i]].x;
    const double dy = (double)(float)p[hull[j]].y - (double)(float)p[hull[i]].y;
    EdgeVec e;
    e.vx = (float)dx; e.vy = (float)dy;
    e.inv = (float)(1.0 / sqrt(dx * dx + dy * dy));
    return e;
}

// out = cx, cy, w, h, angle(deg) as cv2.minAreaRect returns them (angle in [-90, 0)).
YSMR_HD void min_area_rect(const Pt16 *p, const uint16_t *hull, int n, float *out)
{
    float cx, cy, w, h;
    double ang;
    if (n == 1) {
        cx = (float)p[hull[0]].x; cy = (float)p[hull[0]].y; w = 0.f; h = 0.f; ang = 0.0;
    } else if (n == 2) {
        const float x0 = (float)p[hull[0]].x, y0 = (float)p[hull[0]].y;
        const float x1 = (float)p[hull[1]].x, y1 = (float)p[hull[1]].y;
        cx = (x0 + x1) * 0.5f; cy = (y0 + y1) * 0.5f;
        const double dx = (double)x1 - (double)x0, dy = (double)y1 - (double)y0;
        w = (float)sqrt(dx * dx + dy * dy); h = 0.f;
        ang = atan2(dy, dx);
    } else {
        int left = 0, bottom = 0, right = 0, top = 0;
        float lx, rx, ty, by;
        lx = rx = (float)p[hull[0]].x; ty = by = (float)p[hull[0]].y;
        for (int i = 0; i < n; ++i) {
            const float px = (float)p[hull[i]].x, py = (float)p[hull[i]].y;
            if (px < lx) { lx = px; left = i; }
            if (px > rx) { rx = px; right = i; }
            if (py > ty) { ty = py; top = i; }
            if (py < by) { by = py; bottom = i; }
        }
        // hull orientation from the first non-zero cross product of consecutive edges (double)
        float orientation = 0.f;
        {
            EdgeVec e = hull_edge(p, hull, n, n - 1);
            double ax = e.vx, ay = e.vy;
            for (int i = 0; i < n; ++i) {
                e = hull_edge(p, hull, n, i);
                const double bx = e.vx, byy = e.vy;
                const double cvx = ax * byy - ay * bx;
                if (cvx != 0) { orientation = cvx > 0 ? 1.f : -1.f; break; }
                ax = bx; ay = byy;
            }
        }
        float a = orientation, b = 0.f;
        int seq[4] = {bottom, right, top, left};
        EdgeVec ev[4];
        for (int i = 0; i < 4; ++i) ev[i] = hull_edge(p, hull, n, seq[i]);
        float minarea = 3.402823466e+38f;
        int best_left = 0, best_bottom = 0;
        float best_a = 0.f, best_b = 0.f, best_w = 0.f, best_h = 0.f;
        for (int k = 0; k < n; ++k) {
            float dp0 = a * ev[0].vx + b * ev[0].vy;
            float dp1 = -b * ev[1].vx + a * ev[1].vy;
            float dp2 = -a * ev[2].vx - b * ev[2].vy;
            float dp3 = b * ev[3].vx - a * ev[3].vy;
            float maxcos = dp0 * ev[0].inv;
            int m = 0;
            float c1 = dp1 * ev[1].inv; if (c1 > maxcos) { m = 1; maxcos = c1; }
            float c2 = dp2 * ev[2].inv; if (c2 > maxcos) { m = 2; maxcos = c2; }
            float c3 = dp3 * ev[3].inv; if (c3 > maxcos) { m = 3; maxcos = c3; }
            const EdgeVec lead = ev[m];
            const float ux = lead.vx * lead.inv, uy = lead.vy * lead.inv;
            if (m == 0) { a = ux; b = uy; }
            else if (m == 1) { a = uy; b = -ux; }
            else if (m == 2) { a = -ux; b = -uy; }
            else { a = -uy; b = ux; }
            seq[m] = seq[m] + 1 == n ? 0 : seq[m] + 1;
            ev[m] = hull_edge(p, hull, n, seq[m]);
            float dx = (float)p[hull[seq[1]]].x - (float)p[hull[seq[3]]].x;
            float dy = (float)p[hull[seq[1]]].y - (float)p[hull[seq[3]]].y;
            const float width = dx * a + dy * b;
            dx = (float)p[hull[seq[2]]].x - (float)p[hull[seq[0]]].x;
            dy = (float)p[hull[seq[2]]].y - (float)p[hull[seq[0]]].y;
            const float height = -dx * b + dy * a;
            const float area = width * height;
            if (area <= minarea) {
                minarea = area;
                best_left = seq[3]; best_bottom = seq[0];
                best_a = a; best_b = b; best_w = width; best_h = height;
            }
        }
        const float A1 = best_a, B1 = best_b, A2 = -best_b, B2 = best_a;
        const float plx = (float)p[hull[best_left]].x, ply = (float)p[hull[best_left]].y;
        const float pbx = (float)p[hull[best_bottom]].x, pby = (float)p[hull[best_bottom]].y;
        const float C1 = A1 * plx + ply * B1;
        const float C2 = A2 * pbx + pby * B2;
        const float idet = 1.f / (A1 * B2 - A2 * B1);
        const float px = (C1 * B2 - C2 * B1) * idet;
        const float py = (A1 * C2 - A2 * C1) * idet;
        const float o1x = A1 * best_w, o1y = B1 * best_w;
        const float o2x = A2 * best_h, o2y = B2 * best_h;
        cx = px + (o1x + o2x) * 0.5f;
        cy = py + (o1y + o2y) * 0.5f;
        w = (float)sqrt((double)o1x * o1x + (double)o1y * o1y);
        h = (float)sqrt((double)o2x * o2x + (double)o2y * o2y);
        ang = atan2((double)o1y, (double)o1x);
    }
    ang = ang * 180.0 / 3.1415926535897932384626433832795;
    while (ang >= 0) { ang -= 90.0; float t = w; w = h; h = t; }
    while (ang < -90) { ang += 90.0; float t = w; w = h; h = t; }
    out[0] = cx; out[1] = cy; out[2] = w; out[3] = h; out[4] = (float)ang;
}

// Whole K3 for one blob.  Returns the number of contour vertices; when it exceeds `cap` nothing is written to out.
YSMR_HD int blob_rect(const BitImage &img, int x0, int y0, Pt16 *pts, uint16_t *ord, uint16_t *stack, uint16_t *hull,
                      int cap, float *out)
{
    const int n = trace_outer_border(img, x0, y0, pts, cap);
    if (n > cap) return n;
    const int nh = convex_hull(pts, n, ord, stack, hull);
    min_area_rect(pts, hull, nh, out);
    return n;
}

}  // namespace ysmr
