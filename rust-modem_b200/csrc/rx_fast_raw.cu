/* rx_fast_raw.cu -- the fast sps-8 RX kernel fed from the real-valued wire of src/bin/demodulate.rs:29-43 (f32 or i16
 * rows, the PLL lock's samples skipped) with a phase offset per frame (Demodulator::lock_phase, demodulator.rs:32-36,
 * leaves one in every frame's PLL): rx_fast_kernel<..., PF = 7 / 8> (rx_fast.cuh, "RAW").  The frame-invariant phase(n)
 * is parked in tensor memory, every frame adds its offset and evaluates glibc's cosf / sinf while loading.  64-tap
 * low-pass, exact MACs; everything else (OQPSK, other tap counts, odd strides) stays with the generic kernel. */
#include "launch.h"
#include "rx_fast.cuh"

namespace mg {
bool rx_fast_raw_supported(uint32_t n_taps, uint32_t fmt) { return n_taps == 64 && (fmt == 1 || fmt == 2); }
uint64_t rx_fast_raw_tiles(uint64_t K) { return (K + 64 * 4 - 1) / (64 * 4); }
cudaError_t rx_fast_raw_launch(const RxArgs& a, const float* h_taps, cudaStream_t stream)
{
    const bool odd = (a.delay & 1u) != 0;
    if (a.rx_fmt == 1) {
        if (odd) return rx_fast_launch_t<64, 0, false, false, 64, 8, 4, 7, 64>(a, h_taps, stream);
        return rx_fast_launch_t<64, 1, false, false, 64, 8, 4, 7, 64>(a, h_taps, stream);
    }
    if (odd) return rx_fast_launch_t<64, 0, false, false, 64, 8, 4, 8, 64>(a, h_taps, stream);
    return rx_fast_launch_t<64, 1, false, false, 64, 8, 4, 8, 64>(a, h_taps, stream);
}
} /* namespace mg */
