// micro-benchmark: issue throughput of the instructions the K1 front-end is made of (one SM and whole chip).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; informational, numbers quoted in DESIGN.md
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 4096

enum Kind { K_FFMA_REG, K_FFMA_IMM, K_FFMA2, K_FADD2, K_PRMT, K_LOP3, K_IADD3, K_MIX_FFMA2_PRMT, K_MIX_FFMAIMM_PRMT, K_LDS128, K_SHFL,
            K_MIX_FFMA2_FFMAIMM, K_IMAD, K_MIX_FFMA2_IADD, K_COUNT };
static const char *names[] = {"FFMA r,r,r", "FFMA r,imm,r", "FFMA2", "FADD2", "PRMT", "LOP3", "IADD3", "FFMA2+PRMT (1:1)", "FFMAimm+PRMT (1:1)",
                              "LDS.128", "SHFL", "FFMA2+FFMAimm (1:1)", "IMAD", "FFMA2+IADD3 (1:1)"};
static const int per_iter[] = {ILP, ILP, ILP, ILP, ILP, ILP, ILP, 2 * ILP, 2 * ILP, ILP, ILP, 2 * ILP, ILP, 2 * ILP};

template <int KIND>
__global__ void __launch_bounds__(1024) k(float *out, long long *cycles, float b, float c, unsigned sel)
{
    __shared__ __align__(16) float sm[1024 * 4 + 64];
    float a[ILP], a2[ILP];
    unsigned u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x * 1e-3f + i; a2[i] = a[i] + 0.5f; u[i] = threadIdx.x * 2654435761u + i; }
    for (int i = threadIdx.x; i < 1024 * 4 + 64; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == K_FFMA_REG) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (KIND == K_FFMA_IMM || KIND == K_MIX_FFMAIMM_PRMT || KIND == K_MIX_FFMA2_FFMAIMM)
                asm volatile("fma.rn.f32 %0, %0, 0f3F7FFF00, %1;" : "+f"(a[i]) : "f"(c));
            if (KIND == K_FFMA2 || KIND == K_MIX_FFMA2_PRMT || KIND == K_MIX_FFMA2_FFMAIMM || KIND == K_MIX_FFMA2_IADD)
                asm volatile("{ .reg .b64 x, y, z;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %2};\n mov.b64 z, {%3, %3};\n"
                             " fma.rn.f32x2 x, x, y, z;\n mov.b64 {%0, %1}, x; }" : "+f"(a[i]), "+f"(a2[i]) : "f"(b), "f"(c));
            if (KIND == K_FADD2)
                asm volatile("{ .reg .b64 x, z;\n mov.b64 x, {%0, %1};\n mov.b64 z, {%2, %2};\n"
                             " add.rn.f32x2 x, x, z;\n mov.b64 {%0, %1}, x; }" : "+f"(a[i]), "+f"(a2[i]) : "f"(c));
            if (KIND == K_PRMT || KIND == K_MIX_FFMA2_PRMT || KIND == K_MIX_FFMAIMM_PRMT)
                asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(u[i]) : "r"(sel));
            if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(u[i]) : "r"(sel));
            if (KIND == K_IADD3 || KIND == K_MIX_FFMA2_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(sel));
            if (KIND == K_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(sel));
            if (KIND == K_LDS128) {
                float4 v;
                const unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[4 * ((threadIdx.x + i + (u[i] & 1)) & 1023)]);
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
                a[i] = v.x; a2[i] = v.w;
            }
            if (KIND == K_SHFL) u[i] = __shfl_xor_sync(0xffffffffu, u[i], 1);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + a2[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(int threads)
{
    float *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(float) * 1024 * 148); cudaMalloc(&cyc, 8 * 148);
    k<KIND><<<1, threads>>>(out, cyc, 0.99999f, 1e-7f, 0x3210u);
    k<KIND><<<1, threads>>>(out, cyc, 0.99999f, 1e-7f, 0x3210u);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr = (double)ITERS * per_iter[KIND] * (threads / 32);
    printf("%-22s warps/SM=%2d : %.3f warp-instr/clk/SM (%.2f clk per warp-instr per SMSP)\n", names[KIND], threads / 32, warp_instr / (double)h,
           4.0 * (double)h / warp_instr);
    cudaFree(out); cudaFree(cyc);
}

template <int KIND> void both() { run<KIND>(128); run<KIND>(512); run<KIND>(1024); }

int main()
{
    both<K_FFMA_REG>(); both<K_FFMA_IMM>(); both<K_FFMA2>(); both<K_FADD2>(); both<K_PRMT>(); both<K_LOP3>(); both<K_IADD3>(); both<K_IMAD>();
    both<K_MIX_FFMA2_PRMT>(); both<K_MIX_FFMAIMM_PRMT>(); both<K_MIX_FFMA2_FFMAIMM>(); both<K_MIX_FFMA2_IADD>(); both<K_LDS128>(); both<K_SHFL>();
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
