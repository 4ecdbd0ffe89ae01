// micro-benchmark 2 (round 2): issue rates of the instructions the fused front-end is built from, with the operand
// forms the kernel uses (accumulate form d = k * x + d, k immediate), and the ALU-pipe / FMA-pipe mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu ; informational, numbers quoted in DESIGN.md
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 2048

enum Kind {
    K_FFMA_IMM, K_FFMA_REG, K_HFMA2_IMM, K_HFMA2_REG, K_HADD2, K_HSET2, K_HMNMX2, K_IDP4A, K_IDP2A, K_PRMT, K_LOP3, K_SHF, K_IADD3, K_VIMNMX,
    K_VOTE, K_I2F, K_F2I, K_MIX_HFMA2_PRMT, K_MIX_HFMA2_IDP4A, K_MIX_IDP4A_PRMT, K_MIX_FFMA_IDP4A, K_MIX_HFMA2_FFMA, K_MIX_3WAY,
    K_MIX_HFMA2_PRMT_2_1, K_LDS128, K_STS128, K_LDS32, K_COUNT
};
static const char *names[] = {
    "FFMA d=k_imm*x+d", "FFMA d=k_reg*x+d", "HFMA2 d=k_imm*x+d", "HFMA2 d=k_reg*x+d", "HADD2", "HSET2 (set.gt.f16x2)", "HMNMX2", "IDP.4A", "IDP.2A",
    "PRMT", "LOP3", "SHF", "IADD3", "VIMNMX (max.s32)", "VOTE.ballot", "I2F.U8->F32(cvt)", "F2I", "HFMA2imm+PRMT 1:1", "HFMA2imm+IDP4A 1:1",
    "IDP4A+PRMT 1:1", "FFMAimm+IDP4A 1:1", "HFMA2imm+FFMAimm 1:1", "HFMA2imm+PRMT+IDP4A 1:1:1", "HFMA2imm+PRMT 2:1", "LDS.128", "STS.128", "LDS.32"};
static const int per_iter[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 1, 1, 1};

template <int KIND>
__global__ void __launch_bounds__(1024) k(float *out, long long *cycles, float b, float c, unsigned sel, unsigned hb)
{
    __shared__ __align__(16) float sm[1024 * 4 + 64];
    float a[ILP], x[ILP];
    unsigned u[ILP], v[ILP], h[ILP], hx[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = threadIdx.x * 1e-3f + i; x[i] = a[i] * 0.5f + 1.f; u[i] = threadIdx.x * 2654435761u + i; v[i] = u[i] ^ 0x55aa33ccu;
        h[i] = 0x3c003c00u + i + threadIdx.x; hx[i] = 0x38003800u + 3 * i + threadIdx.x;
    }
    for (int i = threadIdx.x; i < 1024 * 4 + 64; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + 16 * threadIdx.x;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            constexpr bool hf_imm = KIND == K_HFMA2_IMM || KIND == K_MIX_HFMA2_PRMT || KIND == K_MIX_HFMA2_IDP4A || KIND == K_MIX_HFMA2_FFMA ||
                                    KIND == K_MIX_3WAY || KIND == K_MIX_HFMA2_PRMT_2_1;
            constexpr bool ff_imm = KIND == K_FFMA_IMM || KIND == K_MIX_FFMA_IDP4A || KIND == K_MIX_HFMA2_FFMA;
            constexpr bool prmt = KIND == K_PRMT || KIND == K_MIX_HFMA2_PRMT || KIND == K_MIX_IDP4A_PRMT || KIND == K_MIX_3WAY || KIND == K_MIX_HFMA2_PRMT_2_1;
            constexpr bool dp4 = KIND == K_IDP4A || KIND == K_MIX_HFMA2_IDP4A || KIND == K_MIX_IDP4A_PRMT || KIND == K_MIX_FFMA_IDP4A || KIND == K_MIX_3WAY;
            if (ff_imm) asm volatile("fma.rn.f32 %0, %1, 0f3E4D0000, %0;" : "+f"(a[i]) : "f"(x[i]));
            if (KIND == K_FFMA_REG) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(x[i]), "f"(b));
            if (hf_imm) asm volatile("{ .reg .b32 kk; mov.b32 kk, 0x32663266; fma.rn.f16x2 %0, %1, kk, %0; }" : "+r"(h[i]) : "r"(hx[i]));
            if (KIND == K_MIX_HFMA2_PRMT_2_1) asm volatile("{ .reg .b32 kk; mov.b32 kk, 0x2e662e66; fma.rn.f16x2 %0, %1, kk, %0; }" : "+r"(hx[i]) : "r"(h[i]));
            if (KIND == K_HFMA2_REG) asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(h[i]) : "r"(hx[i]), "r"(hb));
            if (KIND == K_HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hx[i]));
            if (KIND == K_HSET2) asm volatile("set.gt.u32.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hx[i]));
            if (KIND == K_HMNMX2) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hx[i]));
            if (dp4) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(v[i]), "r"(sel));
            if (KIND == K_IDP2A) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(v[i]), "r"(sel));
            if (prmt) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(v[i]) : "r"(sel));
            if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(sel), "r"(v[i]));
            if (KIND == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(u[i]) : "r"(v[i]));
            if (KIND == K_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[i]));
            if (KIND == K_VIMNMX) asm volatile("max.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[i]));
            if (KIND == K_VOTE) asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "+r"(u[i]));
            if (KIND == K_I2F) asm volatile("{ .reg .b32 t; and.b32 t, %1, 0xff; cvt.rn.f32.u32 %0, t; }" : "=f"(a[i]) : "r"(u[i]));
            if (KIND == K_F2I) asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(u[i]) : "f"(a[i]));
            if (KIND == K_LDS128) {
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(sbase + 16 * i));
                a[i] += q.x; x[i] += q.w;
            }
            if (KIND == K_LDS32) {
                float q;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q) : "r"(sbase + 16 * i));
                a[i] += q;
            }
            if (KIND == K_STS128)
                asm volatile("st.shared.v4.f32 [%4], {%0, %1, %2, %3};" ::"f"(a[i]), "f"(x[i]), "f"(a[i]), "f"(x[i]), "r"(sbase + 16 * i));
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + x[i] + (float)u[i] + (float)v[i] + (float)h[i] + (float)hx[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(int threads)
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(float) * 1024 * 148); cudaMalloc(&cyc, 8 * 148);
    k<KIND><<<1, threads>>>(out, cyc, 0.99999f, 1e-7f, 0x01020304u, 0x32663266u);
    k<KIND><<<1, threads>>>(out, cyc, 0.99999f, 1e-7f, 0x01020304u, 0x32663266u);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr = (double)ITERS * ILP * per_iter[KIND] * (threads / 32);
    printf("%-28s warps/SM=%2d : %.3f warp-instr/clk/SM (%.2f clk per warp-instr per SMSP)\n", names[KIND], threads / 32, warp_instr / (double)h,
           4.0 * (double)h / warp_instr);
    cudaFree(out); cudaFree(cyc);
}

template <int KIND> void both() { run<KIND>(512); run<KIND>(1024); }
template <int K0, int K1> struct Seq { static void go() { both<K0>(); Seq<K0 + 1, K1>::go(); } };
template <int K1> struct Seq<K1, K1> { static void go() {} };

int main()
{
    Seq<0, K_COUNT>::go();
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
