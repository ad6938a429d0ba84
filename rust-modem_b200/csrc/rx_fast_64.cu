/* rx_fast_64.cu -- instantiations of the fast RX kernel for the reference's 64-tap low-pass
 * (src/bin/demodulate.rs:82-147). */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
cudaError_t rx_fast_launch_64(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    return rx_fast_dispatch<64, RX_DEFAULT_THREADS, RX_DEFAULT_MINB, RX_DEFAULT_R>(a, h_taps, fma, tmem, stream);
}
uint64_t rx_fast_tiles_64(uint64_t K)
{
    const uint64_t ts = (uint64_t)RX_DEFAULT_THREADS * RX_DEFAULT_R;
    return (K + ts - 1) / ts;
}
} /* namespace mg */
