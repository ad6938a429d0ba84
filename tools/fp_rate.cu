// fp_rate.cu -- measured FP64 vs FP32 FMA issue rate per SM on the device at hand (decides how expensive the
// binary64 libm port is next to the binary32 FIR work).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_loop(T* out, int iters, T a, T b)
{
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <typename T>
double run(const char* name, int sms, double clk_ghz)
{
    const int blocks = sms * 8, threads = 256, iters = 4096;
    T* d;
    cudaMalloc(&d, sizeof(T) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_loop<T><<<blocks, threads>>>(d, 16, (T)1.0000001, (T)0.5);
    cudaEventRecord(e0);
    fma_loop<T><<<blocks, threads>>>(d, iters, (T)1.0000001, (T)0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * threads * iters * 8;
    const double per_sm_clk = fmas / (ms * 1e-3) / sms / (clk_ghz * 1e9);
    printf("%s: %.3f ms, %.2f TFMA/s, %.1f FMA lanes per SM per clock (at %.3f GHz)\n", name, ms, fmas / ms / 1e9, per_sm_clk, clk_ghz);
    cudaFree(d);
    return per_sm_clk;
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate / 1e6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    run<float>("fp32 fma", p.multiProcessorCount, ghz);
    run<double>("fp64 fma", p.multiProcessorCount, ghz);
    return 0;
}
