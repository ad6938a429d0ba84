/* loop_fused_64.cu -- the fused loopback kernel: rx_fast_kernel with TXF (rx_fast.cuh) for the headline shape
 * (QPSK, rectangular hold, 8 samples per symbol, the reference's 64-tap low-pass of src/bin/demodulate.rs:82-147,
 * odd decision delay, exact MACs, no noise).  One launch replaces tx_rect_fast_kernel + rx_fast_kernel. */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
uint64_t loop_fused_tile_symbols_64() { return (uint64_t)RX_DEFAULT_THREADS * RX_DEFAULT_R; }
/* samples the tiles of a frame reach: the last tile ends at tiles*TS*8 + delay + OFF - 7 (OFF = 0, odd delay) */
bool loop_fused_supported_64(const RxArgs& a)
{
    using C = RxFastCfg<64, 0, RX_DEFAULT_THREADS, RX_DEFAULT_R>;
    if (!(a.delay & 1u) || a.delay + 1 > 8u * C::NB) return false; /* the first tile must start at or before sample 0 */
    const uint64_t tiles = (a.K + C::TS - 1) / C::TS;
    return tiles * C::TS * 8 + a.delay - 7 >= a.L; /* ... and the last one must reach the frame's end */
}
cudaError_t loop_fused_launch_64(const RxArgs& a, const float* h_taps, cudaStream_t stream)
{
    return rx_fast_launch_t<64, 0, false, false, RX_DEFAULT_THREADS, RX_DEFAULT_MINB, RX_DEFAULT_R, RX_DEFAULT_PF, RX_DEFAULT_TMC, true>(a, h_taps, stream);
}
} /* namespace mg */
