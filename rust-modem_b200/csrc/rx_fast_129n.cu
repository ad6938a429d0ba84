/* rx_fast_129n.cu -- the fast RX kernel for the 129-tap matched filter with AWGN added while loading (BASELINE
 * config 4).  The noisy variants carry the Philox / Box-Muller state on top of the FIR's registers: they get the roomier
 * 8-CTA shape (the 10-CTA shape of the noise-free kernel spills there). */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
cudaError_t rx_fast_launch_129n(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    return rx_fast_dispatch_noise<129, 64, 8, 4, RX_DEFAULT_PF, 64>(a, h_taps, fma, tmem, stream);
}
} /* namespace mg */
