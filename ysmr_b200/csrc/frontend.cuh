// Parameters and launchers of the K1 front-end kernels (frontend.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ysmr {

struct FrontParams {
    const uint8_t *frames;       // n_frames frames, H*W*channels bytes each
    int64_t frame_stride;        // bytes between frames
    int n_frames, h, w, ww, channels;
    int t_mask, t_marker;        // d = blurred - mean;  BINARY: d > t ; BINARY_INV: d <= t  (== !(d > t))
    int inverted;                // THRESH_BINARY_INV (dark bacteria)
    const int32_t *scalar_thr;   // mean/std mode: per-frame threshold on the blurred image (NULL otherwise)
    uint32_t *mask_bits;         // [n_frames][h][ww]
    uint32_t *marker_bits;       // [n_frames][h][ww] or NULL (single threshold)
    float k[11];                 // cv2.getGaussianKernel(11, 0, CV_32F)
    int row_tail_from;           // w - w % 4 : first column of OpenCV's scalar tail of the row filter
    int col_tail_from;           // w - w % 8 : first column of the scalar tail of the column filter
    // generation-3 front-end intermediates (owned by the context)
    uint8_t *plane;              // blurred planes: [n_frames] x plane_stride bytes, (h + 10) rows of `pitch` bytes
    int64_t plane_stride;
    int pitch;                   // 16 + round_up(w, 128)
    uint8_t *decisions;          // decision bytes: [n_frames] x dec_stride bytes, h rows of dec_pitch bytes
    int64_t dec_stride;
    int dec_pitch;               // 32 * ceil(w / 128)
    // optional byte-image dumps [n_frames][h][w] (tile kernel only)
    uint8_t *dbg_grey, *dbg_blurred, *dbg_mean;
};

// fused.cu: the generation-4 front-end (one kernel, bound-and-refine)
bool fused_frontend_supported(const FrontParams &p);
cudaError_t fused_frontend_init();                                           // per device, from ysmr_create
cudaError_t launch_fused_frontend(const FrontParams &p, cudaStream_t st);    // 1 launch

cudaError_t launch_frontend_tile(const FrontParams &p, cudaStream_t st);
cudaError_t launch_blur_prepass(const FrontParams &p, cudaStream_t st);      // K1a + margins: 2 launches
cudaError_t launch_gauss_decide(const FrontParams &p, cudaStream_t st, cudaStream_t st_tail, int *n_launched);   // K1b
cudaError_t launch_pack_masks(const FrontParams &p, cudaStream_t st);        // K1c: 1 launch
cudaError_t launch_unpack_bits(const uint32_t *bits, uint8_t *bytes, int64_t rows, int w, int ww, cudaStream_t st);
cudaError_t launch_frame_moments(const uint8_t *frames, int64_t stride, int n_frames, int h, int w, int channels,
                                 unsigned long long *sums, cudaStream_t st);
cudaError_t launch_moving_threshold(const unsigned long long *sums, int n_frames, int h, int w, int white_on_dark,
                                    int signed_offset, int first_frame, int window, double *values, int32_t *thr,
                                    cudaStream_t st);

}  // namespace ysmr
