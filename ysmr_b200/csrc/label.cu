// Device side of K2 (labelling, label.cuh) and K3 (per-blob geometry, geometry.cuh).
#include "kernels.cuh"
#include "label.cuh"
#include "geometry.cuh"

namespace ysmr {

#ifndef LABEL_MIN_CTAS
#define LABEL_MIN_CTAS 4
#endif

// ---- CTA policy for label_frame ------------------------------------------------------------------------------------
struct DevCta {
    uint32_t *warp_sums;   // shared, [33]
    uint32_t *fast;        // dynamic shared memory: parent, seed, kparent [fast_runs each], ext [fast_runs + 1], gparent [fast_runs + h + 2]
    uint32_t fast_runs;
    // Pointer chasing in parent[] / gparent[] dominates the later phases; in shared memory each hop costs ~30 cycles
    // instead of an L2 round trip.  Frames with more runs than fit keep the global scratch.
    __device__ void relocate(LabelFrame &f, uint32_t n_runs) const
    {
        if (n_runs > fast_runs) return;
        uint32_t *q = fast;
        f.parent = q; q += fast_runs;
        f.seed = q; q += fast_runs;
        f.kparent = q; q += fast_runs;
        f.ext = q; q += fast_runs + 1;
        f.gparent = q;
    }
    __device__ int tid() const { return threadIdx.x; }
    __device__ int nthr() const { return blockDim.x; }
    __device__ void sync() const { __syncthreads(); }
    __device__ uint32_t atomic_min(uint32_t *p, uint32_t v) const { return atomicMin(p, v); }
    __device__ void atomic_add(uint32_t *p, uint32_t v) const { atomicAdd(p, v); }
    __device__ void atomic_and(uint32_t *p, uint32_t m) const { atomicAnd(p, m); }
    __device__ void atomic_or_i32(int32_t *p, int32_t v) const { atomicOr(p, v); }

    // In-place exclusive prefix sum of a[0..n) by the whole CTA; returns the total to every thread.
    // Each thread owns a contiguous slice; slice sums are scanned with shuffles + one shared array.
    __device__ uint32_t exclusive_scan(uint32_t *a, int n) const
    {
        const int t = threadIdx.x, nt = blockDim.x;
        const int per = (n + nt - 1) / nt;
        const int lo = min(t * per, n), hi = min(lo + per, n);
        __syncthreads();                                   // producers of a[] are done
        uint32_t sum = 0;
        for (int i = lo; i < hi; ++i) sum += a[i];
        uint32_t incl = sum;
        const int lane = t & 31, warp = t >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int nw = (nt + 31) >> 5;
            uint32_t v = lane < nw ? warp_sums[lane] : 0u;
            uint32_t inc2 = v;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += u;
            }
            if (lane < nw) warp_sums[lane] = inc2 - v;     // exclusive warp offsets
            if (lane == 31) warp_sums[32] = inc2;          // grand total
        }
        __syncthreads();
        uint32_t run = warp_sums[warp] + incl - sum;
        const uint32_t total = warp_sums[32];
        for (int i = lo; i < hi; ++i) { const uint32_t v = a[i]; a[i] = run; run += v; }
        __syncthreads();
        return total;
    }
};

// One CTA per frame slot; CTAs stride over the frames of the batch and reuse their private scratch.
__global__ void __launch_bounds__(LABEL_THREADS, LABEL_MIN_CTAS) label_kernel(LabelLaunch L)
{
    __shared__ uint32_t warp_sums[33];
    extern __shared__ uint32_t fast_mem[];
    DevCta cta{warp_sums, fast_mem, (uint32_t)L.fast_runs};
    const int slot = blockIdx.x;
    uint8_t *base = L.scratch + (size_t)slot * L.scratch_stride;
    for (int f = blockIdx.x; f < L.n_frames; f += gridDim.x) {
        LabelFrame fr;
        fr.h = L.h; fr.w = L.w; fr.ww = L.ww; fr.max_runs = L.max_runs; fr.max_blobs = L.max_blobs;
        fr.mode_propagate = L.mode_propagate;
        fr.img = L.img_bits + (size_t)f * L.h * L.ww;
        fr.seedimg = L.seed_bits ? L.seed_bits + (size_t)f * L.h * L.ww : nullptr;
        // carve the scratch (all sub-arrays 16-byte aligned by construction of the sizes in label_scratch_bytes)
        uint8_t *q = base;
        auto take = [&](size_t bytes) { uint8_t *r = q; q += (bytes + 15) & ~(size_t)15; return r; };
        const size_t R = (size_t)L.max_runs;
        fr.row_start = (uint32_t *)take(4 * (size_t)(L.h + 1));
        uint32_t *krs = (uint32_t *)take(4 * (size_t)(L.h + 1));
        fr.rx0 = (uint16_t *)take(2 * R); fr.rx1 = (uint16_t *)take(2 * R); fr.ry = (uint16_t *)take(2 * R);
        uint16_t *kx0 = (uint16_t *)take(2 * R), *kx1 = (uint16_t *)take(2 * R), *ky = (uint16_t *)take(2 * R);
        fr.parent = (uint32_t *)take(4 * R); fr.seed = (uint32_t *)take(4 * R);
        fr.kparent = (uint32_t *)take(4 * R); fr.ext = (uint32_t *)take(4 * (R + 1));
        fr.gparent = (uint32_t *)take(4 * (R + (size_t)L.h + 2));
        if (L.mode_propagate) { fr.krow_start = krs; fr.kx0 = kx0; fr.kx1 = kx1; fr.ky = ky; }
        else { fr.krow_start = fr.row_start; fr.kx0 = fr.rx0; fr.kx1 = fr.rx1; fr.ky = fr.ry; }
        fr.blob_count = L.blob_count + f;
        fr.first_xy = L.first_xy + (size_t)f * L.max_blobs;
        fr.status = L.status;
        fr.counts = L.counts + (size_t)f * 4;
        if (threadIdx.x < 4) fr.counts[threadIdx.x] = 0;
        __syncthreads();
        label_frame(cta, fr);
        __syncthreads();
        // hand the blobs of this frame to the geometry kernel
        const int nb = *fr.blob_count;
        __shared__ uint32_t work_base;
        if (threadIdx.x == 0) {
            work_base = nb ? atomicAdd(L.work_count, (uint32_t)nb) : 0u;
            if ((fr.counts[0] > (uint32_t)L.max_runs) || (fr.counts[2] > (uint32_t)L.max_blobs)) atomicMin(L.first_bad, L.first_frame + f);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < nb; k += blockDim.x) L.work[work_base + k] = ((uint32_t)f << 16) | (uint32_t)k;
        __syncthreads();
    }
}

size_t label_scratch_bytes(int h, int max_runs)
{
    auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
    const size_t R = (size_t)max_runs;
    return 2 * al(4 * (size_t)(h + 1)) + 6 * al(2 * R) + 3 * al(4 * R) + al(4 * (R + 1)) + al(4 * (R + (size_t)h + 2));
}

// Shared-memory budget per CTA for the union-find arrays: LABEL_MIN_CTAS CTAs of this size still fit one SM.
constexpr size_t LABEL_FAST_BYTES = 45 * 1024;

int label_fast_runs(int h)
{
    const long words = (long)(LABEL_FAST_BYTES / 4) - (long)h - 3;
    const long r = words / 5;
    return r >= 256 ? (int)r : 0;
}

cudaError_t launch_label(const LabelLaunch &L0, int grid, cudaStream_t st)
{
    LabelLaunch L = L0;
    L.fast_runs = label_fast_runs(L.h);
    const size_t smem = L.fast_runs ? 4 * ((size_t)5 * L.fast_runs + (size_t)L.h + 3) : 0;
    label_kernel<<<grid, LABEL_THREADS, smem, st>>>(L);
    return cudaGetLastError();
}

// ---- K3 --------------------------------------------------------------------------------------------------------------
constexpr int GEO_SMALL_CAP = 64;

__global__ void __launch_bounds__(128) geometry_kernel(GeoLaunch G)
{
    const uint32_t total = *G.work_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t item = G.work[i];
        const int f = item >> 16, k = item & 0xFFFF;
        const uint32_t fxy = G.first_xy[(size_t)f * G.max_blobs + k];
        BitImage img{G.img_bits + (size_t)f * G.h * G.ww, G.h, G.w, G.ww};
        Pt16 pts[GEO_SMALL_CAP];
        uint16_t ord[GEO_SMALL_CAP + 4], stk[GEO_SMALL_CAP + 4], hull[GEO_SMALL_CAP + 4];
        float out[5];
        const int n = blob_rect(img, (int)(fxy & 0xFFFF), (int)(fxy >> 16), pts, ord, stk, hull, GEO_SMALL_CAP, out);
        float *dst = G.blobs + ((size_t)f * G.max_blobs + k) * 5;
        if (n <= GEO_SMALL_CAP) {
#pragma unroll
            for (int j = 0; j < 5; ++j) dst[j] = out[j];
            if (G.first_xy_dbg) { G.first_xy_dbg[((size_t)f * G.max_blobs + k) * 2] = fxy & 0xFFFF; G.first_xy_dbg[((size_t)f * G.max_blobs + k) * 2 + 1] = fxy >> 16; }
        } else {
            // long contour: defer to the big-blob kernel with scratch from the pool
            const uint32_t slot = atomicAdd(G.big_count, 1u);
            if (slot < (uint32_t)G.big_cap) { G.big_items[2 * slot] = item; G.big_items[2 * slot + 1] = (uint32_t)n; }
            else { atomicOr(G.status, 4); atomicMin(G.first_bad, G.first_frame + f); dst[0] = dst[1] = dst[2] = dst[3] = dst[4] = 0.f; }
            if (G.first_xy_dbg) { G.first_xy_dbg[((size_t)f * G.max_blobs + k) * 2] = fxy & 0xFFFF; G.first_xy_dbg[((size_t)f * G.max_blobs + k) * 2 + 1] = fxy >> 16; }
        }
    }
}

// Contours longer than GEO_SMALL_CAP vertices: one thread per blob, arrays carved from a global pool.
__global__ void __launch_bounds__(64) geometry_big_kernel(GeoLaunch G)
{
    uint32_t total = *G.big_count;
    if (total > (uint32_t)G.big_cap) total = G.big_cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t item = G.big_items[2 * i];
        const int n = (int)G.big_items[2 * i + 1];
        const int f = item >> 16, k = item & 0xFFFF;
        float *dst = G.blobs + ((size_t)f * G.max_blobs + k) * 5;
        const size_t need = ((size_t)n * 4 + 3 * ((size_t)n + 4) * 2 + 15) & ~(size_t)15;
        bool ok = n <= 65000;
        unsigned long long off = 0;
        if (ok) {
            off = atomicAdd(G.pool_used, (unsigned long long)need);
            ok = off + need <= G.pool_bytes;
        }
        if (!ok) {
            atomicOr(G.status, 4); atomicMin(G.first_bad, G.first_frame + f);
            dst[0] = dst[1] = dst[2] = dst[3] = dst[4] = 0.f;
            continue;
        }
        uint8_t *q = G.pool + off;
        Pt16 *pts = (Pt16 *)q; q += (size_t)n * 4;
        uint16_t *ord = (uint16_t *)q; q += ((size_t)n + 4) * 2;
        uint16_t *stk = (uint16_t *)q; q += ((size_t)n + 4) * 2;
        uint16_t *hull = (uint16_t *)q;
        const uint32_t fxy = G.first_xy[(size_t)f * G.max_blobs + k];
        BitImage img{G.img_bits + (size_t)f * G.h * G.ww, G.h, G.w, G.ww};
        float out[5];
        blob_rect(img, (int)(fxy & 0xFFFF), (int)(fxy >> 16), pts, ord, stk, hull, n, out);
        for (int j = 0; j < 5; ++j) dst[j] = out[j];
    }
}

cudaError_t launch_geometry(const GeoLaunch &G, cudaStream_t st)
{
    geometry_kernel<<<148 * 4, 128, 0, st>>>(G);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    geometry_big_kernel<<<148, 64, 0, st>>>(G);
    return cudaGetLastError();
}

}  // namespace ysmr
