// Shared definitions of the ysmr_b200 CUDA library (sm_100a).
//
// The per-item logic of the labelling, geometry and linking kernels is written as plain inline functions marked
// YSMR_HD so that the very same source can be compiled for the host by tests/host_emul/ (test infrastructure: it
// lets the CPU-only test suite exercise the kernel logic against cv2/scipy without a GPU).  The product library
// libysmr_b200.so only ever runs them on the device.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define YSMR_HD __host__ __device__ __forceinline__
#define YSMR_D __device__ __forceinline__
#define YSMR_HD_NOINLINE static __host__ __device__ __noinline__      // rare or bulky leaves: keep the callers' loop bodies small
#else
#define YSMR_HD inline
#define YSMR_D inline
#define YSMR_HD_NOINLINE inline
struct float2 { float x, y; };           // (host emulation only: the vector types come from cuda_runtime.h otherwise)
#endif

namespace ysmr {

// Packed binary image: 1 bit per pixel, bit (x & 31) of word [y * ww + (x >> 5)], bits beyond the width are zero.
struct BitImage {
    const uint32_t *bits;
    int h, w, ww;
    YSMR_HD bool at(int x, int y) const
    {
        if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return false;
        return (bits[(int64_t)y * ww + (x >> 5)] >> (x & 31)) & 1u;
    }
};

YSMR_HD int words_per_row(int w) { return (w + 31) >> 5; }

// ---- float helpers with explicit rounding (the library is compiled with -fmad=false; host emulation with
// ---- -ffp-contract=off), so a*b+c below is two roundings everywhere and fmaf/fma is the only fused form.
#if defined(__CUDA_ARCH__)
YSMR_D float fmaf_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
#else
}  // namespace ysmr
#include <math.h>
namespace ysmr {
inline float fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
#endif

}  // namespace ysmr
