// check_sincos_dev.cu -- the DEVICE evaluation of libm_f32.h's cosf / sinf against this machine's glibc, exhaustively:
// every binary32 with |y| < 120 (both signs) through mg_sincosf_nco -- the routine behind every NCO angle of the form
// phase + offset -- and through the general mg_sincosf.  The GPU reduces each run of 2^20 consecutive bit patterns to a
// 64-bit checksum of the (sin, cos) result bits; the host computes the same checksums with glibc's sinf / cosf (OpenMP) and
// compares.  (tools/check_libm.c proves the ALGORITHM equals glibc when run on the host; this proves the device's binary64
// operations and conversions give the same bits.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -fmad=false -Xcompiler -fopenmp -I rust-modem_b200/csrc -o tools/bin/check_sincos_dev tools/check_sincos_dev.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "libm_f32.h"

constexpr uint32_t RUN = 1u << 20;
__host__ __device__ inline uint64_t fold(uint32_t sb, uint32_t cb, uint32_t u) { return (((uint64_t)sb << 32) | cb) * 0x9E3779B97F4A7C15ull + u; }

template <int WHICH>
__global__ void sums(uint32_t top, uint32_t neg, unsigned long long* out)
{
    const uint32_t base = blockIdx.x * RUN;
    unsigned long long acc = 0;
    for (uint32_t i = threadIdx.x; i < RUN; i += blockDim.x) {
        const uint32_t u = base + i;
        if (u >= top) break;
        const float y = __uint_as_float(u | (neg << 31));
        float s, c;
        if (WHICH == 0) mg_sincosf_nco(y, &s, &c);
        else mg_sincosf(y, &s, &c);
        acc += fold(__float_as_uint(s), __float_as_uint(c), u);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + blockIdx.x, acc);
}
__global__ void atan_sums(unsigned long long* out)
{
    const uint64_t base = (uint64_t)blockIdx.x * RUN;
    unsigned long long acc = 0;
    for (uint32_t i = threadIdx.x; i < RUN; i += blockDim.x) {
        const uint32_t u = (uint32_t)(base + i);
        const float r = mg_atanf(__uint_as_float(u));
        uint32_t rb = __float_as_uint(r);
        if (r != r) rb = 0x7fc00000u; /* one NaN */
        acc += fold(rb, 0u, u);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + blockIdx.x, acc);
}
/* atanf (the PLL's atan2f is built on it): ALL 2^32 binary32 inputs on the device against glibc's atanf */
static long check_atan()
{
    const uint32_t runs = 1u << 12;
    unsigned long long* d;
    cudaMalloc(&d, runs * 8);
    cudaMemset(d, 0, runs * 8);
    atan_sums<<<runs, 256>>>(d);
    std::vector<unsigned long long> got(runs), want(runs, 0);
    if (cudaMemcpy(got.data(), d, runs * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return -1; }
#pragma omp parallel for schedule(dynamic)
    for (uint32_t r = 0; r < runs; ++r) {
        unsigned long long acc = 0;
        for (uint32_t i = 0; i < RUN; ++i) {
            const uint32_t u = r * RUN + i;
            float y;
            memcpy(&y, &u, 4);
            const float a = atanf(y);
            uint32_t ab;
            memcpy(&ab, &a, 4);
            if (a != a) ab = 0x7fc00000u;
            acc += fold(ab, 0u, u);
        }
        want[r] = acc;
    }
    long b = 0;
    for (uint32_t r = 0; r < runs; ++r) b += got[r] != want[r];
    printf("mg_atanf (device): all 4294967296 binary32 inputs in %u runs, %ld runs differ from glibc\n", runs, b);
    cudaFree(d);
    return b;
}

int main()
{
    float lim = 120.0f;
    uint32_t top;
    memcpy(&top, &lim, 4);
    top += 4096; /* a little beyond 120: the dispatching wrapper's general branch */
    const uint32_t runs = (top + RUN - 1) / RUN;
    unsigned long long* d;
    cudaMalloc(&d, runs * 8);
    long bad = 0, total = 0;
    for (int which = 0; which < 2; ++which)
        for (uint32_t neg = 0; neg < 2; ++neg) {
            cudaMemset(d, 0, runs * 8);
            if (which == 0) sums<0><<<runs, 256>>>(top, neg, d);
            else sums<1><<<runs, 256>>>(top, neg, d);
            std::vector<unsigned long long> got(runs), want(runs, 0);
            if (cudaMemcpy(got.data(), d, runs * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 2; }
#pragma omp parallel for schedule(dynamic)
            for (uint32_t r = 0; r < runs; ++r) {
                unsigned long long acc = 0;
                for (uint32_t i = 0; i < RUN; ++i) {
                    const uint32_t u = r * RUN + i;
                    if (u >= top) break;
                    const uint32_t yb = u | (neg << 31);
                    float y;
                    memcpy(&y, &yb, 4);
                    const float s = sinf(y), c = cosf(y);
                    uint32_t sb, cb;
                    memcpy(&sb, &s, 4);
                    memcpy(&cb, &c, 4);
                    acc += fold(sb, cb, u);
                }
                want[r] = acc;
            }
            long b = 0;
            for (uint32_t r = 0; r < runs; ++r) b += got[r] != want[r];
            printf("%s, %s inputs: %u values in %u runs, %ld runs differ from glibc\n", which == 0 ? "mg_sincosf_nco (device)" : "mg_sincosf (device)",
                   neg ? "negative" : "positive", top, runs, b);
            bad += b;
            total += top;
        }
    printf("%ld device evaluations against this machine's sinf / cosf: %s\n", total, bad ? "MISMATCH" : "BIT-EXACT");
    const long ba = check_atan();
    return (bad || ba) ? 1 : 0;
}
