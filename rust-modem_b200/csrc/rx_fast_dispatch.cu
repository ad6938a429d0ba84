/* rx_fast_dispatch.cu -- tap-count dispatch over the rx_fast_<NT>.cu translation units. */
#include "launch.h"

namespace mg {
cudaError_t rx_fast_launch_64(const RxArgs&, const float*, bool, bool, cudaStream_t);
cudaError_t rx_fast_launch_129(const RxArgs&, const float*, bool, bool, cudaStream_t);
uint64_t rx_fast_tiles_64(uint64_t);
uint64_t rx_fast_tiles_129(uint64_t);

/* the no-TMEM instantiations exist for the exact MAC only */
bool rx_fast_supported(uint32_t n_taps, bool fma, bool tmem) { return (n_taps == 64 || n_taps == 129) && (tmem || !fma); }
uint64_t rx_fast_tiles(uint32_t n_taps, uint64_t K) { return n_taps == 64 ? rx_fast_tiles_64(K) : rx_fast_tiles_129(K); }
cudaError_t rx_fast_launch(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    return a.n_taps == 64 ? rx_fast_launch_64(a, h_taps, fma, tmem, stream) : rx_fast_launch_129(a, h_taps, fma, tmem, stream);
}
} /* namespace mg */
