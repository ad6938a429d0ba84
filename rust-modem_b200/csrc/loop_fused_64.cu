/* loop_fused_64.cu -- the fused loopback kernel: rx_fast_kernel with TXF (rx_fast.cuh) for the headline shape
 * (one 4-point table, rectangular hold, 8 samples per symbol, the reference's 64-tap low-pass of
 * src/bin/demodulate.rs:82-147, odd decision delay, exact MACs, no noise).  One launch replaces
 * tx_rect_fast_kernel + rx_fast_kernel.
 *
 * CTA shape (measured on B200 at C2, profiles/r02_fused_variants.txt): 128 threads x 4 symbols per thread, 4 CTAs per SM
 * (16 warps), 64 TMEM columns per CTA -- 256 of the SM's 512 columns, on all 128 lanes.  64-thread CTAs at 8 per SM use
 * all 512 columns for the same speed; more CTAs with 32 columns and fewer registers are 5-10 % slower; 8 symbols per
 * thread (32 % less FIR shared traffic) needs 150-170 registers, i.e. 12 warps per SM: 0.62 ms at best (FUSED_R etc. below
 * rebuild that shape). */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
#ifndef FUSED_R
#define FUSED_R 4
#define FUSED_THREADS 128
#define FUSED_MINB 4
#define FUSED_TMC 64
#endif
constexpr int kFusedThreads = FUSED_THREADS, kFusedMinB = FUSED_MINB, kFusedR = FUSED_R, kFusedTmc = FUSED_TMC;
uint64_t loop_fused_tile_symbols_64() { return (uint64_t)kFusedThreads * kFusedR; }
/* samples the tiles of a frame reach: the last tile ends at tiles*TS*8 + delay + OFF - 7 (OFF = 0, odd delay) */
bool loop_fused_supported_64(const RxArgs& a)
{
    using C = RxFastCfg<64, 0, kFusedThreads, kFusedR>;
    if (!(a.delay & 1u) || a.delay + 1 > 8u * C::NB) return false; /* the first tile must start at or before sample 0 */
    const uint64_t tiles = (a.K + C::TS - 1) / C::TS;
    return tiles * C::TS * 8 + a.delay - 7 >= a.L; /* ... and the last one must reach the frame's end */
}
cudaError_t loop_fused_launch_64(const RxArgs& a, const float* h_taps, bool tmem, cudaStream_t stream)
{
    if (!tmem) return rx_fast_launch_t<64, 0, false, false, kFusedThreads, kFusedMinB, kFusedR, 3, 0, true>(a, h_taps, stream);
    return rx_fast_launch_t<64, 0, false, false, kFusedThreads, kFusedMinB, kFusedR, 3, kFusedTmc, true>(a, h_taps, stream);
}
} /* namespace mg */
