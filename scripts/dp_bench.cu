// micro-benchmark: FP64 / FP32 FMA throughput and latency on one SM (informational; used for DESIGN.md)
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T *out, int iters, long long *cycles)
{
    T a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
    const T b = (T)1.0000001, c = (T)1e-7;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = a[i] * b + c;
    long long t1 = clock64();
    T s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <typename T, int ILP>
void run(const char *name, int threads)
{
    T *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(T) * 2048); cudaMalloc(&cyc, 8 * 8);
    const int iters = 2000;
    fma_kernel<T, ILP><<<1, threads>>>(out, iters, cyc);
    fma_kernel<T, ILP><<<1, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr = (double)iters * ILP * (threads / 32);
    printf("%-6s threads=%4d ILP=%d : %.2f cycles per warp-FMA per SM  (%.1f lanes/clk/SM), %.1f cyc per dependent step\n", name, threads, ILP,
           (double)h / warp_instr, 32.0 * warp_instr / (double)h, (double)h / iters);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<double, 1>("f64", 32); run<double, 4>("f64", 32); run<double, 4>("f64", 256); run<double, 8>("f64", 1024);
    run<float, 1>("f32", 32); run<float, 4>("f32", 256); run<float, 8>("f32", 1024);
    return 0;
}
