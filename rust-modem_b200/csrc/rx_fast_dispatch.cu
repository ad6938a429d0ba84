/* rx_fast_dispatch.cu -- tap-count dispatch over the rx_fast_<NT>.cu translation units. */
#include "launch.h"

namespace mg {
cudaError_t rx_fast_launch_64(const RxArgs&, const float*, bool, int, cudaStream_t);
cudaError_t rx_fast_launch_129(const RxArgs&, const float*, bool, int, cudaStream_t);
uint64_t rx_fast_tiles_64(uint64_t, int);
uint64_t rx_fast_tiles_129(uint64_t, int);
cudaError_t rx_ws_launch_64(const RxArgs&, const float*, bool, int, cudaStream_t);
uint64_t rx_ws_tiles_64(uint64_t, int);

bool rx_fast_supported(uint32_t n_taps) { return n_taps == 64 || n_taps == 129; }
uint64_t rx_fast_tiles(uint32_t n_taps, uint64_t K, int variant)
{
    if (n_taps == 64 && variant == 10) return rx_ws_tiles_64(K, variant);
    return n_taps == 64 ? rx_fast_tiles_64(K, variant) : rx_fast_tiles_129(K, variant);
}
cudaError_t rx_fast_launch(const RxArgs& a, const float* h_taps, bool fma, int variant, cudaStream_t stream)
{
    if (a.n_taps == 64 && variant == 10 && (a.delay & 1u) && !fma && a.nz.sigma == 0.0f)
        return rx_ws_launch_64(a, h_taps, fma, variant, stream);
    return a.n_taps == 64 ? rx_fast_launch_64(a, h_taps, fma, variant, stream)
                          : rx_fast_launch_129(a, h_taps, fma, variant, stream);
}
} /* namespace mg */
