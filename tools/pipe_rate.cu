// pipe_rate.cu -- measured issue rates of the binary32 forms the FIR can be built from, per SM and clock, on the device
// at hand: scalar FFMA / FMUL+FADD, packed FMUL2 / FFMA2 with register and with uniform (constant-bank) operands, and
// mixes of them.  Decides what the exact two-rounding MAC of fir.rs:23 costs on the FMA pipe.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bin/pipe_rate pipe_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fmul(float a, float b) { float d; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

struct P { float h[16]; float one[2]; };
// MODE: 0 scalar ffma (reg,reg,reg)  1 scalar ffma (reg,const,reg)  2 FMUL2(reg,const)+FFMA2(reg,const,reg) = the exact MAC
//       3 FFMA2(reg,const,reg) only  4 FMUL2(reg,const) only  5 FFMA2(reg,reg,reg)  6 scalar fmul(reg,const)+fadd(reg,reg) x2 rails
//       7 FMUL2 + 2 scalar fadd (per MAC pair)   8 scalar fmul x2 + FFMA2(acc,1,p)
template <int MODE>
__global__ void k(float* out, const float* in, int iters, const __grid_constant__ P p)
{
    constexpr int N = 8;
    u64 acc[N], v[N];
    float r0 = in[threadIdx.x], r1 = in[threadIdx.x + 32];
    for (int i = 0; i < N; ++i) { acc[i] = pk(threadIdx.x + i, i); v[i] = pk(in[i], in[i + 8]); }
    const u64 one = pk(p.one[0], p.one[1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const u64 hh = pk(p.h[t], p.h[t]);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 0) { float2 a = *reinterpret_cast<float2*>(&acc[i]); a.x = ffma(a.x, r0, r1); a.y = ffma(a.y, r1, r0); acc[i] = pk(a.x, a.y); }
                if (MODE == 1) { float2 a = *reinterpret_cast<float2*>(&acc[i]); a.x = ffma(a.x, p.h[t], r1); a.y = ffma(a.y, p.h[t], r0); acc[i] = pk(a.x, a.y); }
                if (MODE == 2) acc[i] = fma2(acc[i], one, mul2(acc[(i + 1) % N], hh)); /* the multiplicand changes every step: nothing can be hoisted out of the loop */
                if (MODE == 3) acc[i] = fma2(acc[i], hh, v[i]);
                if (MODE == 4) acc[i] = mul2(acc[i], hh);
                if (MODE == 5) acc[i] = fma2(acc[i], v[i], v[(i + 1) % N]);
                if (MODE == 6) { float2 a = *reinterpret_cast<float2*>(&acc[i]); float2 w = *reinterpret_cast<float2*>(&acc[(i + 1) % N]); a.x = fadd(a.x, fmul(w.x, p.h[t])); a.y = fadd(a.y, fmul(w.y, p.h[t])); acc[i] = pk(a.x, a.y); }
                if (MODE == 7) { float2 a = *reinterpret_cast<float2*>(&acc[i]); u64 pr = mul2(acc[(i + 1) % N], hh); float2 w = *reinterpret_cast<float2*>(&pr); a.x = fadd(a.x, w.x); a.y = fadd(a.y, w.y); acc[i] = pk(a.x, a.y); }
                if (MODE == 8) { float2 w = *reinterpret_cast<float2*>(&acc[(i + 1) % N]); acc[i] = fma2(acc[i], one, pk(fmul(w.x, p.h[t]), fmul(w.y, p.h[t]))); }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < N; ++i) { float2 a = *reinterpret_cast<float2*>(&acc[i]); s += a.x + a.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double lane_ops_per_mac_pair, int sms, double ghz, float* d, const float* in, int threads, int bps)
{
    P p; for (int i = 0; i < 16; ++i) p.h[i] = 1.0f + 1e-7f * i; p.one[0] = p.one[1] = 1.0f;
    const int blocks = sms * bps, iters = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, in, 4, p);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, in, iters, p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double pairs = (double)blocks * threads * iters * 16 * 8; // (both-rail) MAC-like steps
    printf("%-58s %2d warps/SM: %.3f ms  %.1f rail-pair steps/SM/clk  (%.1f binary32 lane-ops/SM/clk)\n", name, threads * bps / 32, ms,
           pairs / (ms * 1e-3) / sms / (ghz * 1e9), pairs * lane_ops_per_mac_pair / (ms * 1e-3) / sms / (ghz * 1e9));
}
int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk / 1e6; const int sms = pr.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz nominal\n", pr.name, sms, ghz);
    float *d, *in; cudaMalloc(&d, sizeof(float) * sms * 16 * 1024); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int threads = cfg ? 64 : 256, bps = cfg ? 8 : 8; // 64 warps/SM, then the RX kernel's 16 warps/SM
        run<0>("scalar FFMA reg,reg,reg (x2)", 2, sms, ghz, d, in, threads, bps);
        run<1>("scalar FFMA reg,const,reg (x2)", 2, sms, ghz, d, in, threads, bps);
        run<2>("FMUL2(reg,const) + FFMA2(acc,one,p): the exact MAC", 4, sms, ghz, d, in, threads, bps);
        run<3>("FFMA2 reg,const,reg only (fused MAC)", 2, sms, ghz, d, in, threads, bps);
        run<4>("FMUL2 reg,const only", 2, sms, ghz, d, in, threads, bps);
        run<5>("FFMA2 reg,reg,reg", 2, sms, ghz, d, in, threads, bps);
        run<6>("scalar FMUL + FADD per rail (x2)", 4, sms, ghz, d, in, threads, bps);
        run<7>("FMUL2 + 2 scalar FADD", 4, sms, ghz, d, in, threads, bps);
        run<8>("2 scalar FMUL + FFMA2(acc,one,p)", 4, sms, ghz, d, in, threads, bps);
    }
    return 0;
}
