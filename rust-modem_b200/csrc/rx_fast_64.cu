/* rx_fast_64.cu -- instantiations of the fast RX kernel for the reference's 64-tap low-pass
 * (src/bin/demodulate.rs:82-147). */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
/* tuning variants (MODEM_GPU_RX_VARIANT), only for the headline shape: odd delay, exact MAC, no noise */
struct V { int threads, r; };
static V variant_shape(int variant)
{
    switch (variant) {
    case 4: case 5: case 6: return {128, 4};
    default: return {64, 4};
    }
}
cudaError_t rx_fast_launch_64(const RxArgs& a, const float* h_taps, bool fma, int variant, cudaStream_t stream)
{
    /* tuning variants (MODEM_GPU_RX_VARIANT), headline shape only: odd delay, exact MAC, no noise */
    if (variant && (a.delay & 1u) && !fma && a.nz.sigma == 0.0f) {
        switch (variant) {
        case 1: return rx_fast_launch_t<64, 0, false, false, 64, 10, 4, 3, 32>(a, h_taps, stream);
        case 2: return rx_fast_launch_t<64, 0, false, false, 64, 8, 4, 3, 64>(a, h_taps, stream);
        case 3: return rx_fast_launch_t<64, 0, false, false, 64, 9, 4, 3, 32>(a, h_taps, stream);
        case 4: return rx_fast_launch_t<64, 0, false, false, 128, 5, 4, 3, 64>(a, h_taps, stream);
        case 5: return rx_fast_launch_t<64, 0, false, false, 128, 6, 4, 3, 64>(a, h_taps, stream);
        case 6: return rx_fast_launch_t<64, 0, false, false, 128, 6, 4, 3, 32>(a, h_taps, stream);
        case 7: return rx_fast_launch_t<64, 0, false, false, 64, 11, 4, 3, 32>(a, h_taps, stream);
        case 8: return rx_fast_launch_t<64, 0, false, false, 64, 7, 4, 3, 64>(a, h_taps, stream);
        case 9: return rx_fast_launch_t<64, 0, false, false, 64, 6, 4, 3, 64>(a, h_taps, stream);
        default: break;
        }
    }
    return rx_fast_dispatch<64, RX_DEFAULT_THREADS, RX_DEFAULT_MINB, RX_DEFAULT_R>(a, h_taps, fma, stream);
}
uint64_t rx_fast_tiles_64(uint64_t K, int variant)
{
    const V v = variant ? variant_shape(variant) : V{RX_DEFAULT_THREADS, RX_DEFAULT_R};
    const uint64_t ts = (uint64_t)v.threads * v.r;
    return (K + ts - 1) / ts;
}
} /* namespace mg */
