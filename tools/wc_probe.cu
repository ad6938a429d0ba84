// wc_probe.cu -- does write-combined pinned memory change the host->device rate of the e2e step's copies?
// H2D of 64 MiB from default pinned and from cudaHostAllocWriteCombined memory, alone and beside a D2H of 64 MiB, as 1 and as 16
// transfers per direction (two streams, CUDA events).  build: nvcc -O2 -o tools/bin/wc_probe tools/wc_probe.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
static float run(void* d_in, const void* h_in, void* h_out, const void* d_out, size_t n, int parts, bool both, cudaStream_t s0, cudaStream_t s1)
{
    cudaEvent_t a, b, c;
    cudaEventCreate(&a); cudaEventCreate(&b); cudaEventCreate(&c);
    float best = 1e9f;
    for (int it = 0; it < 8; ++it) {
        cudaDeviceSynchronize();
        cudaEventRecord(a, s0);
        cudaStreamWaitEvent(s1, a, 0);
        for (int p = 0; p < parts; ++p) {
            cudaMemcpyAsync((char*)d_in + p * (n / parts), (const char*)h_in + p * (n / parts), n / parts, cudaMemcpyHostToDevice, s0);
            if (both) cudaMemcpyAsync((char*)h_out + p * (n / parts), (const char*)d_out + p * (n / parts), n / parts, cudaMemcpyDeviceToHost, s1);
        }
        cudaEventRecord(c, s1);
        cudaStreamWaitEvent(s0, c, 0);
        cudaEventRecord(b, s0);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it >= 2 && ms < best) best = ms;
    }
    return best;
}
int main()
{
    const size_t n = 64u << 20;
    void *h_def, *h_wc, *h_out, *d_in, *d_out;
    cudaHostAlloc(&h_def, n, cudaHostAllocDefault);
    cudaHostAlloc(&h_wc, n, cudaHostAllocWriteCombined);
    cudaHostAlloc(&h_out, n, cudaHostAllocDefault);
    cudaMalloc(&d_in, n); cudaMalloc(&d_out, n);
    memset(h_def, 1, n); memset(h_wc, 1, n); memset(h_out, 0, n);
    cudaStream_t s0, s1; cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    for (int parts : {1, 16})
        for (int both = 0; both < 2; ++both) {
            const float td = run(d_in, h_def, h_out, d_out, n, parts, both, s0, s1), tw = run(d_in, h_wc, h_out, d_out, n, parts, both, s0, s1);
            printf("%2d transfer(s) per direction, %s: default pinned %.3f ms (%.1f GB/s)   write-combined %.3f ms (%.1f GB/s)\n", parts,
                   both ? "in + out" : "in only ", td, n / td / 1e6, tw, n / tw / 1e6);
        }
    return 0;
}
