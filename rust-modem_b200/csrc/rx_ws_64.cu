/* rx_ws_64.cu -- instantiations of the warp-specialised fast RX kernel for the reference's
 * 64-tap low-pass (src/bin/demodulate.rs:82-147). */
#include "launch.h"
#include "rx_ws.cuh"

namespace mg {
/* Experimental (MODEM_GPU_RX_VARIANT=10): correct, halves the L1/shared wavefronts, but with ~1.5
 * consumer warps per scheduler the FMUL2->FFMA2 latencies are exposed; slower than rx_fast today. */
cudaError_t rx_ws_launch_64(const RxArgs& a, const float* h_taps, bool fma, int variant, cudaStream_t stream)
{
    (void)variant;
    (void)fma;
    if ((a.delay & 1u) && a.nz.sigma == 0.0f) return rx_ws_launch_t<64, 0, false, false, 4, 2, 4, 3>(a, h_taps, stream);
    return cudaErrorNotSupported;
}
uint64_t rx_ws_tiles_64(uint64_t K, int variant)
{
    (void)variant;
    return (K + 255) / 256;
}
} /* namespace mg */
