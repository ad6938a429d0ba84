/* rx_fast_64n.cu -- the fast RX kernel for the 64-tap low-pass with AWGN added while loading (Philox + Box-Muller in
 * phase A, rx_fast.cuh).  Same CTA shape as the noise-free kernel. */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
cudaError_t rx_fast_launch_64n(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    return rx_fast_dispatch_noise<64, RX_DEFAULT_THREADS, RX_DEFAULT_MINB, RX_DEFAULT_R, RX_DEFAULT_PF, RX_DEFAULT_TMC>(a, h_taps, fma, tmem, stream);
}
} /* namespace mg */
