/*
 * rx_fullrate_fast.cu -- the reference's Demodulator iterator at full rate (demodulator.rs:44-55 with its two
 * FIRFilter::add, fir.rs:18-34): filtered (I, Q) for EVERY input sample, 64-tap low-pass (demodulate.rs:82-147).
 *
 * 2 x 64 exact MACs per sample = 256 FP32 lane-operations: at 128 lanes per SM and clock the kernel is bound by
 * the FP32 pipes at ~145 Gsamples/s (exact) / ~290 (fused), far below the HBM roofline (16 B per sample would
 * allow ~410).  The job is therefore to keep the FMA pipe busy with nothing but MACs:
 *   - one thread computes RO = 8 consecutive outputs with a sliding register window: tap k multiplies the eight
 *     samples n+r-k (r = 0..7); moving to tap k+1 shifts the window by one sample, so per tap there is ONE
 *     uniform tap operand, half a 128-bit shared load and 8 packed MACs (FMUL2 + FFMA2, both rails at once);
 *   - every accumulator sees its taps in ascending order k = 0..63, newest sample first (fir.rs:21-24);
 *   - (vi, vq) = (x cos t, x (-sin t)) pairs are staged once per tile, two samples per 16-byte chunk, one padding
 *     chunk per four so the threads' 128-bit loads (stride 5 chunks) are bank-conflict free;
 *   - NCO values come from the context's table when the phase offset is per call, or are evaluated per sample
 *     when every frame carries its own PLL offset (ChannelView::po_frame); the tile's values stay in registers
 *     while the CTA loops over its frames.
 */
#include "launch.h"

namespace mg {

/* (cos, sin) of phase + per-frame offset as one out-of-line routine (see rx_fast.cuh: inlined copies of the binary64 routine
 * make the kernel stall on instruction fetch) */
static __device__ __noinline__ float2 sincos_nco_ool(float y)
{
    float s, c;
    mg_sincosf_nco(y, &s, &c);
    return make_float2(c, s);
}


template <int NT>
struct FrCfg {
    static constexpr int THREADS = 128;
    static constexpr int RO = 8;                        /* outputs per thread */
    static constexpr int TILE = THREADS * RO;           /* outputs per CTA tile */
    static constexpr int HALO = (NT - 1 + 1) & ~1;      /* history samples staged before the tile (even) */
    static constexpr int NS = TILE + HALO;              /* staged samples */
    static constexpr int NCH = NS / 2;                  /* 16-byte chunks */
    static constexpr int PCH = NCH + NCH / 4 + 1;       /* with one padding chunk per 4 */
    static constexpr int PER = (NS + THREADS - 1) / THREADS; /* staged samples per thread */
};

template <int NT, bool FMA, bool PFP>
__global__ void __launch_bounds__(128, 4)
    rx_fullrate_fast_kernel(const __grid_constant__ RxArgs a, const __grid_constant__ TapsParam<NT> taps)
{
    using C = FrCfg<NT>;
    __shared__ __align__(16) float2 s_v[2 * C::PCH];
    const int tid = threadIdx.x;
    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const u64 t0 = (u64)blockIdx.x * C::TILE;                     /* first output sample of the tile */
    const long long nb = (long long)t0 - C::HALO;                 /* sample index of staged element 0 */
    const f32x2 one = pk2(taps.one.x, taps.one.y);

    /* frame-invariant part of the NCO for this thread's staged samples: (cos, sin) when the phase offset is the
     * same for every frame of the CTA, else the phase itself (the offset is added per frame, demodulator.rs:50) */
    float nc[C::PER], ns[C::PER];
    bool ok[C::PER];
    const float w = chan_w(a.ch, f0);
    const float2* tab = PFP ? nullptr : chan_table(a.ch, f0);
    const float po0 = PFP ? 0.0f : chan_po(a.ch, f0);
#pragma unroll
    for (int i = 0; i < C::PER; ++i) {
        const int j = tid + i * C::THREADS;
        const long long n = nb + j;
        ok[i] = j < C::NS && n >= 0 && (u64)n < a.L;
        nc[i] = ns[i] = 0.0f;
        if (ok[i]) {
            if (PFP) {
                nc[i] = nco_phase(w, a.sample0 + (u64)n);
            } else if (tab) {
                const float2 t = __ldg(tab + n);
                nc[i] = t.x;
                ns[i] = t.y;
            } else {
                mg_sincosf_nco(__fadd_rn(nco_phase(w, a.sample0 + (u64)n), po0), &ns[i], &nc[i]);
            }
        }
    }

    float* s_f = reinterpret_cast<float*>(s_v);
    /* the tile's input samples of the NEXT frame are fetched into registers before the FIR of the current one:
     * the long FIR (1024 packed operations per thread) hides the global latency (measured: the unpipelined form
     * spent 5.6 stall cycles per issued instruction on the load scoreboard) */
    float xr[C::PER];
    auto fetch = [&](u64 f) {
#pragma unroll
        for (int i = 0; i < C::PER; ++i) xr[i] = ok[i] ? rx_sample(a, f, (u64)(nb + tid + i * C::THREADS)) : 0.0f;
    };
    if (f0 < f1) fetch(f0);
    for (u64 f = f0; f < f1; ++f) {
        const float po = PFP ? chan_po(a.ch, f) : 0.0f;
        __syncthreads(); /* previous frame's readers are done */
#pragma unroll
        for (int i = 0; i < C::PER; ++i) {
            const int j = tid + i * C::THREADS;
            if (j < C::NS) {
                float vi = 0.0f, vq = 0.0f; /* zero history before the frame (fir.rs:13) and past its end */
                if (ok[i]) {
                    const float x = xr[i];
                    float c = nc[i], s = ns[i];
                    if (PFP) { const float2 t = sincos_nco_ool(__fadd_rn(nc[i], po)); c = t.x; s = t.y; }
                    vi = __fmul_rn(x, c);  /* demodulator.rs:53 */
                    vq = __fmul_rn(x, -s); /* demodulator.rs:54 */
                }
                const int ch = j >> 1, pos = ch + (ch >> 2); /* one padding chunk per four */
                *reinterpret_cast<float2*>(s_f + 4 * pos + 2 * (j & 1)) = make_float2(vi, vq);
            }
        }
        __syncthreads();
        if (f + 1 < f1) fetch(f + 1);

        /* outputs n = t0 + 8*tid + r, r = 0..7: staged index of output r is HALO + 8*tid + r; tap k reads staged
         * index HALO + 8*tid + r - k.  Window registers win[r] hold the sample for output r at the current tap. */
        const int base = C::HALO + 8 * tid; /* staged index of output 0 */
        auto ld_chunk = [&](int chunk) -> ulonglong2 { /* two packed samples (even, odd staged index) */
            const int pos = chunk + (chunk >> 2);
            return *reinterpret_cast<const ulonglong2*>(s_f + 4 * pos);
        };
        f32x2 win[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const ulonglong2 t = ld_chunk(base / 2 + q); /* base is even: chunk-aligned */
            win[2 * q] = t.x;
            win[2 * q + 1] = t.y;
        }
        f32x2 acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0ull; /* (+0.0f, +0.0f) */
        ulonglong2 older = make_ulonglong2(0ull, 0ull);
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const f32x2 hh = pk2(taps.hh[k].x, taps.hh[k].y);
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = mac2<FMA>(acc[r], win[r], hh, one);
            if (k + 1 < NT) {
                /* shift the window one sample towards the past: the new element is staged index base - (k + 1) */
                if (((k + 1) & 1) == 1) older = ld_chunk(base / 2 - (k + 2) / 2); /* chunk holding base-k-2, base-k-1 */
#pragma unroll
                for (int r = 7; r > 0; --r) win[r] = win[r - 1];
                win[0] = ((k + 1) & 1) ? older.y : older.x;
            }
        }
        float4* out = reinterpret_cast<float4*>(a.filt + f * a.L + t0 + 8 * tid);
        const u64 n_out0 = t0 + 8 * (u64)tid;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 y0 = unpk2(acc[2 * q]), y1 = unpk2(acc[2 * q + 1]);
            const float4 v = make_float4(__fmul_rn(a.rx_gain, y0.x), __fmul_rn(a.rx_gain, y0.y), __fmul_rn(a.rx_gain, y1.x),
                                         __fmul_rn(a.rx_gain, y1.y)); /* demodulator.rs:53-54: 2.0 * lp.add(..) */
            if (n_out0 + 2 * q + 1 < a.L) {
                __stcs(out + q, v);
            } else if (n_out0 + 2 * q < a.L) {
                __stcs(a.filt + f * a.L + n_out0 + 2 * q, make_float2(v.x, v.y));
            }
        }
    }
}

bool rx_fullrate_fast_supported(uint32_t n_taps) { return n_taps == 64; }
uint64_t rx_fullrate_fast_tiles(uint64_t L) { return (L + FrCfg<64>::TILE - 1) / FrCfg<64>::TILE; }

cudaError_t rx_fullrate_fast_launch(const RxArgs& a, const float* h_taps, bool fma, bool per_frame_po, cudaStream_t stream)
{
    dim3 grid((unsigned)rx_fullrate_fast_tiles(a.L), (unsigned)((a.F + a.frames_per_block - 1) / a.frames_per_block));
    const TapsParam<64> tp = make_taps_param<64>(h_taps);
    if (per_frame_po) {
        if (fma) rx_fullrate_fast_kernel<64, true, true><<<grid, 128, 0, stream>>>(a, tp);
        else rx_fullrate_fast_kernel<64, false, true><<<grid, 128, 0, stream>>>(a, tp);
    } else {
        if (fma) rx_fullrate_fast_kernel<64, true, false><<<grid, 128, 0, stream>>>(a, tp);
        else rx_fullrate_fast_kernel<64, false, false><<<grid, 128, 0, stream>>>(a, tp);
    }
    return cudaGetLastError();
}

} /* namespace mg */
