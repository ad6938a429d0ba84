/* rx_fast_64.cu -- noise-free instantiations of the fast RX kernel for the reference's 64-tap low-pass
 * (src/bin/demodulate.rs:82-147); the noisy ones are in rx_fast_64n.cu. */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
cudaError_t rx_fast_launch_64n(const RxArgs&, const float*, bool, bool, cudaStream_t);
cudaError_t rx_fast_launch_64(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    if (a.nz.sigma != 0.0f) return rx_fast_launch_64n(a, h_taps, fma, tmem, stream);
    return rx_fast_dispatch_clean<64, RX_DEFAULT_THREADS, RX_DEFAULT_MINB, RX_DEFAULT_R, RX_DEFAULT_PF, RX_DEFAULT_TMC>(a, h_taps, fma, tmem, stream);
}
uint64_t rx_fast_tiles_64(uint64_t K)
{
    const uint64_t ts = (uint64_t)RX_DEFAULT_THREADS * RX_DEFAULT_R;
    return (K + ts - 1) / ts;
}
} /* namespace mg */
