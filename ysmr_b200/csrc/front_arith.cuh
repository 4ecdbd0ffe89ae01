// Device arithmetic shared by the front-end kernels (frontend.cu: K1a / K1b / K1c and the debug tile kernel; fused.cu: the
// fused bound-and-refine kernel).  Everything here is bit-exact with OpenCV 4.13 (oracle/c_stages.c).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ysmr {

__device__ __forceinline__ int reflect101(int p, int n)
{
    // one reflection is exact for -n < p < 2n-1 (all positions whose value is ever used: halos of at most 6 with
    // n >= 16); positions further out only occur in tiles hanging over the image edge and are clamped to stay in bounds
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p < 0 ? 0 : (p >= n ? n - 1 : p);
}

__device__ __forceinline__ int clampi(int p, int n) { return p < 0 ? 0 : (p >= n ? n - 1 : p); }

__device__ __forceinline__ uint32_t luma(uint32_t b, uint32_t g, uint32_t r)
{
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

constexpr float KG0 = 0.00881223008f, KG1 = 0.0271435864f, KG2 = 0.0651140586f, KG3 = 0.121649072f, KG4 = 0.176998362f,
                KG5 = 0.200565413f;

template <int C>
__device__ __forceinline__ uint32_t grey_px(const uint8_t *frame, int w, int y, int x)
{
    const uint8_t *p = frame + ((int64_t)y * w + x) * C;
    if (C == 3) return luma(p[0], p[1], p[2]);
    return p[0];
}

template <bool TAIL>
__device__ __forceinline__ float gauss_row(const float *a, bool tail)
{
    // a[0..10] = blurred[x-5 .. x+5]
    float acc = __fmul_rn(KG0, a[0]);
    if (TAIL && tail) {
        acc = __fadd_rn(acc, __fmul_rn(KG1, a[1])); acc = __fadd_rn(acc, __fmul_rn(KG2, a[2]));
        acc = __fadd_rn(acc, __fmul_rn(KG3, a[3])); acc = __fadd_rn(acc, __fmul_rn(KG4, a[4]));
        acc = __fadd_rn(acc, __fmul_rn(KG5, a[5])); acc = __fadd_rn(acc, __fmul_rn(KG4, a[6]));
        acc = __fadd_rn(acc, __fmul_rn(KG3, a[7])); acc = __fadd_rn(acc, __fmul_rn(KG2, a[8]));
        acc = __fmaf_rn(KG1, a[9], acc); acc = __fmaf_rn(KG0, a[10], acc);
        return acc;
    }
    acc = __fmaf_rn(KG1, a[1], acc); acc = __fmaf_rn(KG2, a[2], acc); acc = __fmaf_rn(KG3, a[3], acc);
    acc = __fmaf_rn(KG4, a[4], acc); acc = __fmaf_rn(KG5, a[5], acc); acc = __fmaf_rn(KG4, a[6], acc);
    acc = __fmaf_rn(KG3, a[7], acc); acc = __fmaf_rn(KG2, a[8], acc); acc = __fmaf_rn(KG1, a[9], acc);
    acc = __fmaf_rn(KG0, a[10], acc);
    return acc;
}

// Blackwell packed FP32: one instruction, two IEEE-rounded results (SASS FFMA2 / FMUL2 / FADD2).  Element-wise identical
// to the scalar operations, so the bit-exactness argument is unchanged; it halves the issue slots of the column pass.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b)
{
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd; }\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0, %1}, rd; }\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// column pass for two adjacent pixels at once
template <bool TAIL>
__device__ __forceinline__ float2 gauss_col2(float2 c, float2 m1, float2 p1, float2 m2, float2 p2, float2 m3, float2 p3, float2 m4,
                                             float2 p4, float2 m5, float2 p5, bool tail)
{
    const float2 k5 = make_float2(KG5, KG5), k4 = make_float2(KG4, KG4), k3 = make_float2(KG3, KG3), k2 = make_float2(KG2, KG2),
                 k1 = make_float2(KG1, KG1), k0 = make_float2(KG0, KG0);
    float2 acc = fmul2(k5, c);
    const float2 s1 = fadd2(m1, p1), s2 = fadd2(m2, p2), s3 = fadd2(m3, p3), s4 = fadd2(m4, p4), s5 = fadd2(m5, p5);
    if (TAIL && tail) {
        acc = fadd2(acc, fmul2(k4, s1)); acc = fadd2(acc, fmul2(k3, s2)); acc = fadd2(acc, fmul2(k2, s3));
        acc = fadd2(acc, fmul2(k1, s4)); acc = fadd2(acc, fmul2(k0, s5));
        return acc;
    }
    acc = ffma2(k4, s1, acc); acc = ffma2(k3, s2, acc); acc = ffma2(k2, s3, acc); acc = ffma2(k1, s4, acc); acc = ffma2(k0, s5, acc);
    return acc;
}

template <bool TAIL>
__device__ __forceinline__ float gauss_col(float c, float m1, float p1, float m2, float p2, float m3, float p3, float m4,
                                           float p4, float m5, float p5, bool tail)
{
    float acc = __fmul_rn(KG5, c);
    const float s1 = __fadd_rn(m1, p1), s2 = __fadd_rn(m2, p2), s3 = __fadd_rn(m3, p3), s4 = __fadd_rn(m4, p4),
                s5 = __fadd_rn(m5, p5);
    if (TAIL && tail) {
        acc = __fadd_rn(acc, __fmul_rn(KG4, s1)); acc = __fadd_rn(acc, __fmul_rn(KG3, s2));
        acc = __fadd_rn(acc, __fmul_rn(KG2, s3)); acc = __fadd_rn(acc, __fmul_rn(KG1, s4));
        acc = __fadd_rn(acc, __fmul_rn(KG0, s5));
        return acc;
    }
    acc = __fmaf_rn(KG4, s1, acc); acc = __fmaf_rn(KG3, s2, acc); acc = __fmaf_rn(KG2, s3, acc);
    acc = __fmaf_rn(KG1, s4, acc); acc = __fmaf_rn(KG0, s5, acc);
    return acc;
}

// Q15 luma g = (3735 B + 19235 G + 9798 R + 16384) >> 15 of a pixel whose three bytes sit anywhere in one or two words, as
// two 16-bit x 8-bit dot products (dp2a) with the weights arranged for the byte position -- no byte shuffling.  All weights
// and the rounding constant are DOUBLED, so the sum is 2 * (...) and g is exactly byte 2 of it: a later byte permute picks
// it up for free instead of a shift per pixel.
constexpr uint32_t LW_BG = 7470u | (38470u << 16), LW_R0 = 19596u, LW_0B = 7470u << 16, LW_GR = 38470u | (19596u << 16);
__device__ __forceinline__ uint32_t lsum_b012(uint32_t v) { return __dp2a_hi(LW_R0, v, __dp2a_lo(LW_BG, v, 32768u)); }
__device__ __forceinline__ uint32_t lsum_b123(uint32_t v) { return __dp2a_hi(LW_GR, v, __dp2a_lo(LW_0B, v, 32768u)); }
__device__ __forceinline__ uint32_t lsum_b3_01(uint32_t v, uint32_t n) { return __dp2a_lo(LW_GR, n, __dp2a_hi(LW_0B, v, 32768u)); }
__device__ __forceinline__ uint32_t lsum_b23_0(uint32_t v, uint32_t n) { return __dp2a_lo(LW_R0, n, __dp2a_hi(LW_BG, v, 32768u)); }

// four BGR pixels (12 bytes = words w0, w1, w2) -> one word of four grey bytes
__device__ __forceinline__ uint32_t grey4_of_bgr(uint32_t w0, uint32_t w1, uint32_t w2)
{
    const uint32_t a = __byte_perm(lsum_b012(w0), lsum_b3_01(w0, w1), 0x0062);     // (g0, g1, -, -)
    const uint32_t b = __byte_perm(lsum_b23_0(w1, w2), lsum_b123(w2), 0x0062);     // (g2, g3, -, -)
    return __byte_perm(a, b, 0x5410);
}

}  // namespace ysmr
