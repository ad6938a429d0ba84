/* rx_fast_129.cu -- noise-free instantiations of the fast RX kernel for the 129-tap root-raised-cosine
 * matched filter (span 16 symbols x 8 samples/symbol + 1); the noisy ones are in rx_fast_129n.cu. */
#include "launch.h"
#include "rx_fast.cuh"

/* CTA shape, measured on B200 at C3 after the round-2 rework of phases A and C (profiles/r02_c3_variants.txt): 64 threads,
 * R = 4, 8 CTAs/SM at 128 registers with 64 TMEM columns = 2.66 ms; 10 CTAs/SM at 96 registers (small spills) with 32
 * columns = 2.73 ms (round 1's choice, then 3.5 % ahead); 128-thread CTAs 2.95-2.99 ms.  The kernel sits at 74 % FMA-pipe
 * utilisation (profiles/r02_c3_rx129_ncu.txt): it is bound by the two rounded operations per tap the reference's MAC
 * consists of, not by memory. */
#ifndef RX129_THREADS
#define RX129_THREADS 64
#define RX129_MINB 8
#define RX129_R 4
#define RX129_TMC 64
#endif
#ifndef RX129_PF
#define RX129_PF RX_DEFAULT_PF
#endif

namespace mg {
cudaError_t rx_fast_launch_129n(const RxArgs&, const float*, bool, bool, cudaStream_t);
cudaError_t rx_fast_launch_129(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    if (a.nz.sigma != 0.0f) return rx_fast_launch_129n(a, h_taps, fma, tmem, stream);
    return rx_fast_dispatch_clean<129, RX129_THREADS, RX129_MINB, RX129_R, RX129_PF, RX129_TMC>(a, h_taps, fma, tmem, stream);
}
uint64_t rx_fast_tiles_129(uint64_t K)
{
    const uint64_t ts = (uint64_t)RX129_THREADS * RX129_R;
    return (K + ts - 1) / ts;
}
} /* namespace mg */
