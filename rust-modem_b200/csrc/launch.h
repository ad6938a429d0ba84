/*
 * launch.h -- host-side launchers of the compile-time specialised kernels (each lives in
 * its own translation unit so the build parallelises).  Every launcher returns false when
 * it has no instantiation for the requested shape; the caller then uses the generic kernel.
 */
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace mg {

struct LaunchGeom {
    uint64_t tiles_x;   /* filled by *_tiles(): CTAs along the sample axis */
};

/* rx_fast_<NT>.cu: sps 8, NT taps */
bool rx_fast_supported(uint32_t n_taps, bool fma, bool tmem);
uint64_t rx_fast_tiles(uint32_t n_taps, uint64_t K);
cudaError_t rx_fast_launch(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream);
/* rx_fast_raw.cu: the same kernel fed from real f32 (fmt 1) / i16 (fmt 2) rows with a phase offset per frame */
bool rx_fast_raw_supported(uint32_t n_taps, uint32_t fmt);
uint64_t rx_fast_raw_tiles(uint64_t K);
cudaError_t rx_fast_raw_launch(const RxArgs& a, const float* h_taps, cudaStream_t stream);
/* fused loopback (loop_fused_64.cu): TX samples made and stored by the demodulating kernel itself */
uint64_t loop_fused_tile_symbols_64();
bool loop_fused_supported_64(const RxArgs& a);
cudaError_t loop_fused_launch_64(const RxArgs& a, const float* h_taps, bool tmem, cudaStream_t stream);

/* rx_dec.cu: tuned decimating RX for any samples-per-symbol count (the reference's default rates: 45), 64 taps */
bool rx_dec_supported(uint32_t n_taps, uint32_t sps);
uint32_t rx_dec_tile_symbols(uint32_t sps);
cudaError_t rx_dec_launch(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream);
/* the same kernel fed from real f32 (fmt 1) / i16 (fmt 2) rows with a phase offset per frame (rx_dec_kernel<..., 2 / 3>) */
bool rx_dec_raw_supported(uint32_t n_taps, uint32_t sps, uint32_t fmt);
cudaError_t rx_dec_raw_launch(const RxArgs& a, const float* h_taps, cudaStream_t stream);
/* the fused loopback for those shapes (rx_dec_kernel<..., TXF>): QPSK-sized table, rectangular hold, sps <= 64 */
bool loop_fused_dec_supported(const RxArgs& a);
cudaError_t loop_fused_dec_launch(const RxArgs& a, const float* h_taps, bool tmem, cudaStream_t stream);

/* rx_fullrate_fast.cu: the full-rate (I,Q) stream, 64 taps; needs L even and a 16-byte aligned filt row base */
bool rx_fullrate_fast_supported(uint32_t n_taps);
uint64_t rx_fullrate_fast_tiles(uint64_t L);
cudaError_t rx_fullrate_fast_launch(const RxArgs& a, const float* h_taps, bool fma, bool per_frame_po, cudaStream_t stream);

/* tx_fast.cu */
bool tx_rect_fast_supported(uint32_t bps);
uint64_t tx_rect_fast_tiles(uint64_t L);
cudaError_t tx_rect_fast_launch(const TxArgs& a, cudaStream_t stream);
bool tx_shaped_fast_supported(uint32_t sps, uint32_t n_taps);
uint64_t tx_shaped_fast_tiles(uint64_t nsym);
cudaError_t tx_shaped_fast_launch(const TxArgs& a, const float* h_taps, bool fma, bool rail_pairs, cudaStream_t stream);

} /* namespace mg */
