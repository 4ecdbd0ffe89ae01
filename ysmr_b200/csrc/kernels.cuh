// Launch descriptors shared between the kernels (label.cu, link.cu) and the C-ABI layer (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ysmr {

constexpr int LABEL_THREADS = 256;
constexpr int LINK_THREADS = 512;    // warps 0-7: one lane per track (fast path), warps 8-15: detection staging + FIR helpers

struct LabelLaunch {
    int n_frames, first_frame, h, w, ww, max_runs, max_blobs, mode_propagate;
    uint32_t *img_bits;          // [n_frames][h][ww]  mask in, findContours image out
    const uint32_t *seed_bits;   // [n_frames][h][ww]  markers (PROPAGATE) or NULL
    uint8_t *scratch;            // grid x scratch_stride bytes
    size_t scratch_stride;
    int32_t *blob_count;         // [n_frames]
    uint32_t *first_xy;          // [n_frames][max_blobs]
    uint32_t *counts;            // [n_frames][4]
    uint32_t *work;              // [n_frames * max_blobs] compact (frame << 16 | blob) list for the geometry kernel
    uint32_t *work_count;        // [1]
    int32_t *status, *first_bad; // [1] each
    int fast_runs;               // set by launch_label: runs per frame whose union-find arrays fit shared memory
};

struct GeoLaunch {
    int h, w, ww, max_blobs, first_frame;
    const uint32_t *img_bits;
    const uint32_t *first_xy;
    const uint32_t *work;
    const uint32_t *work_count;
    float *blobs;                // [n_frames][max_blobs][5]
    int32_t *first_xy_dbg;       // optional [n_frames][max_blobs][2]
    uint32_t *big_items;         // [big_cap][2]
    uint32_t *big_count;         // [1]
    int big_cap;
    uint8_t *pool;               // contour scratch for long contours
    unsigned long long pool_bytes;
    unsigned long long *pool_used;
    int32_t *status, *first_bad;
};

size_t label_scratch_bytes(int h, int max_runs);
cudaError_t launch_label(const LabelLaunch &L, int grid, cudaStream_t st);
cudaError_t launch_geometry(const GeoLaunch &G, cudaStream_t st);

}  // namespace ysmr
