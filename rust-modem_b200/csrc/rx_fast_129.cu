/* rx_fast_129.cu -- instantiations of the fast RX kernel for the 129-tap root-raised-cosine
 * matched filter (span 16 symbols x 8 samples/symbol + 1). */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
cudaError_t rx_fast_launch_129(const RxArgs& a, const float* h_taps, bool fma, int variant, cudaStream_t stream)
{
    (void)variant;
    return rx_fast_dispatch<129, RX_DEFAULT_THREADS, RX_DEFAULT_MINB, RX_DEFAULT_R>(a, h_taps, fma, stream);
}
uint64_t rx_fast_tiles_129(uint64_t K, int variant)
{
    (void)variant;
    return (K + RX_DEFAULT_R * RX_DEFAULT_THREADS - 1) / (RX_DEFAULT_R * RX_DEFAULT_THREADS);
}
} /* namespace mg */
