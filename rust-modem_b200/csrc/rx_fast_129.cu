/* rx_fast_129.cu -- instantiations of the fast RX kernel for the 129-tap root-raised-cosine
 * matched filter (span 16 symbols x 8 samples/symbol + 1). */
#include "launch.h"
#include "rx_fast.cuh"

/* measured on B200 at C3 (profiles/r01_c3_variants.txt): 64 threads, R = 4, 10 CTAs/SM (96 registers, 12 bytes
 * of spill) with 32 TMEM columns beats the 64-tap kernel's shape (8 CTAs/SM at 128 registers) by 3.5 % */
#ifndef RX129_THREADS
#define RX129_THREADS 64
#define RX129_MINB 10
#define RX129_R 4
#define RX129_TMC 32
#endif

namespace mg {
/* tuning variants (MODEM_GPU_RX_VARIANT 21..29), matched-filter shape only: even delay (128), exact MAC, no noise */
struct V129 { int threads, r; };
static V129 variant_shape_129(int variant)
{
    switch (variant) {
    case 21: return {64, 8};
    case 22: return {128, 4};
    case 23: return {64, 4};
    case 24: return {128, 8};
    case 25: return {64, 4};
    case 26: return {64, 8};
    case 27: return {128, 4};
    default: return {RX129_THREADS, RX129_R};
    }
}
cudaError_t rx_fast_launch_129(const RxArgs& a, const float* h_taps, bool fma, int variant, cudaStream_t stream)
{
    if (variant >= 21 && !(a.delay & 1u) && !fma && a.nz.sigma == 0.0f) {
        switch (variant) {
        case 21: return rx_fast_launch_t<129, 1, false, false, 64, 6, 8, 3, 64>(a, h_taps, stream);
        case 22: return rx_fast_launch_t<129, 1, false, false, 128, 4, 4, 3, 64>(a, h_taps, stream);
        case 23: return rx_fast_launch_t<129, 1, false, false, 64, 10, 4, 3, 32>(a, h_taps, stream);
        case 24: return rx_fast_launch_t<129, 1, false, false, 128, 3, 8, 3, 64>(a, h_taps, stream);
        case 25: return rx_fast_launch_t<129, 1, false, false, 64, 8, 4, 3, 0>(a, h_taps, stream);
        case 26: return rx_fast_launch_t<129, 1, false, false, 64, 8, 8, 3, 32>(a, h_taps, stream);
        case 27: return rx_fast_launch_t<129, 1, false, false, 128, 5, 4, 3, 32>(a, h_taps, stream);
        default: break;
        }
    }
    /* the noisy variants carry the Philox / Box-Muller state on top of the FIR's registers: give them the roomier
     * 8-CTA shape (the 10-CTA shape spills 80 bytes there) */
    if (a.nz.sigma != 0.0f) return rx_fast_dispatch<129, 64, 8, 4, RX_DEFAULT_PF, 64>(a, h_taps, fma, stream);
    return rx_fast_dispatch<129, RX129_THREADS, RX129_MINB, RX129_R, RX_DEFAULT_PF, RX129_TMC>(a, h_taps, fma, stream);
}
uint64_t rx_fast_tiles_129(uint64_t K, int variant)
{
    const V129 v = variant_shape_129(variant);
    const uint64_t ts = (uint64_t)v.threads * v.r;
    return (K + ts - 1) / ts;
}
} /* namespace mg */
