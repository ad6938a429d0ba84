/*
 * tx_fast.cu -- compile-time specialised TX kernels (rectangular hold with word-wise bit
 * loads; 129-tap pulse shaping at 8 samples per symbol) and their launchers.
 */
#include <stdlib.h>

#include <type_traits>
#include "launch.h"

namespace mg {

/*
 * Fast rectangular-hold TX: one table, no Q offset, even sps (both samples of a 128-bit
 * store share a symbol), bits rows aligned to BPS bytes.  The frame loop is unrolled by FU
 * with all bit loads of the batch issued first, so FU*U independent loads are in flight
 * per thread instead of one dependent chain per frame.
 */
/* TWO: odd samples-per-symbol (the reference's default rates give 45): the two samples of a store may belong to
 * different symbols, so each gets its own bit word and table lookup */
template <int BPS, bool REAL, bool TWO>
__global__ void __launch_bounds__(kThreads, 4) tx_rect_fast_kernel(const __grid_constant__ TxArgs a)
{
    constexpr int U = 2, FU = 4;
    __shared__ float2 s_lut[1 << BPS];
    for (uint32_t i = threadIdx.x; i < (1u << BPS); i += kThreads) s_lut[i] = a.lut[i];
    __syncthreads();

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);
    const float2* tab = chan_table(a.ch, f0);

    uint32_t koff[U]; /* byte offset of the symbol's bits inside a frame row */
    uint32_t koff1[U]; /* TWO: the same for the pair's second sample */
    uint32_t noff[U]; /* float4 offset of the sample pair inside a frame row */
    bool valid[U];
    float c0[U], s0[U], c1[U], s1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t pair = ((uint32_t)blockIdx.x * U + u) * kThreads + threadIdx.x;
        const u64 n = 2 * (u64)pair;
        valid[u] = n < a.L;
        noff[u] = pair;
        koff[u] = (uint32_t)(n / a.sps) * BPS;
        koff1[u] = TWO ? (uint32_t)((n + 1) / a.sps) * BPS : koff[u];
        if (tab && n + 1 < a.L) { /* n is even and the table row is 16-byte aligned */
            const float4 t = __ldg(reinterpret_cast<const float4*>(tab + n));
            c0[u] = t.x; s0[u] = t.y; c1[u] = t.z; s1[u] = t.w;
        } else {
            mg_sincosf(nco_phase(w, a.sample0 + n), &s0[u], &c0[u]);
            mg_sincosf(nco_phase(w, a.sample0 + n + 1), &s1[u], &c1[u]);
        }
    }

    const uint8_t* pb = a.bits + f0 * a.nbits;
    float4* po = REAL ? nullptr : reinterpret_cast<float4*>(a.tx + f0 * a.L);
    const u64 Lq = a.L / 2; /* float4 per frame row */
    /* REAL: the wire format of src/bin/modulate.rs:128-133 -- `x.modulate().re` only, 4 bytes per sample; rows may
     * start at an odd float (odd preamble length), so the pair is written as two 32-bit stores */
    float* pr = REAL ? a.re + f0 * a.re_stride + a.re_offset : nullptr;
    auto emit = [&](int j, int u, float2 o0, float2 o1) {
        if (REAL) {
            float* q = pr + (u64)j * a.re_stride + 2 * (u64)noff[u];
            __stcs(q, o0.x);
            __stcs(q + 1, o1.x);
        } else {
            __stcs(po + j * Lq + noff[u], make_float4(o0.x, o0.y, o1.x, o1.y));
        }
    };
    u64 f = f0;
    for (; f + FU <= f1; f += FU) {
        uint32_t idx[FU][U], idx1[FU][U];
#pragma unroll
        for (int j = 0; j < FU; ++j)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                idx[j][u] = valid[u] ? load_symbol_word<BPS>(pb + j * a.nbits + koff[u]) : 0u;
                idx1[j][u] = TWO ? (valid[u] ? load_symbol_word<BPS>(pb + j * a.nbits + koff1[u]) : 0u) : idx[j][u];
            }
#pragma unroll
        for (int j = 0; j < FU; ++j)
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float2 bb = s_lut[idx[j][u]];
                const float2 b1 = TWO ? s_lut[idx1[j][u]] : bb;
                const float2 o0 = mix_iq(bb.x, bb.y, c0[u], s0[u]);
                const float2 o1 = mix_iq(b1.x, b1.y, c1[u], s1[u]);
                if (valid[u]) emit(j, u, o0, o1);
            }
        pb += FU * a.nbits;
        if (REAL) pr += FU * a.re_stride;
        else po += FU * Lq;
    }
    for (; f < f1; ++f) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u]) continue;
            const float2 bb = s_lut[load_symbol_word<BPS>(pb + koff[u])];
            const float2 b1 = TWO ? s_lut[load_symbol_word<BPS>(pb + koff1[u])] : bb;
            const float2 o0 = mix_iq(bb.x, bb.y, c0[u], s0[u]);
            const float2 o1 = mix_iq(b1.x, b1.y, c1[u], s1[u]);
            emit(0, u, o0, o1);
        }
        pb += a.nbits;
        if (REAL) pr += a.re_stride;
        else po += Lq;
    }
}

/*
 * Fast pulse-shaped TX for compile-time (SPS, NT): one thread owns one symbol period (SPS consecutive
 * samples) of FB frames at a time.  The taps are kernel-parameter constants: each (h, h) pair is ONE 64-bit
 * uniform-register operand fetched from the constant bank, and it feeds the MACs of all FB frames before the
 * next one is fetched (with one frame per fetch the uniform loads cost more issue slots than the MACs:
 * 84 thread-instructions per sample in profiles/r01, 2/3 of them LDCU/UMOV).  The walk is symbol-major: the
 * J = ceil(NT/SPS) symbols that reach the period are read one at a time from shared memory (one 64-bit
 * conflict-free load per frame) and each feeds the SPS phase accumulators with taps p + j*SPS, so every
 * accumulator still sees its taps in ascending order (fir.rs:21-24; the zero-stuffed terms are exact no-ops).
 * The CTA's 2 KB-per-warp output is transposed through shared memory so global stores are 128-bit coalesced.
 */
template <int SPS, int NT, bool FMA, int FB, int BPSW>
__global__ void __launch_bounds__(kThreads, FB >= 4 ? 2 : 3)
    tx_shaped_fast_kernel(const __grid_constant__ TxArgs a, const __grid_constant__ TapsParam<NT> taps)
{
    static_assert(SPS == 8, "output transpose below is written for 8 samples per symbol");
    constexpr int J = (NT + SPS - 1) / SPS; /* symbols reaching one output */
    constexpr int HALO = J - 1;
    constexpr int ROW = kThreads + HALO;
    __shared__ float2 s_lut[kMaxLut];
    __shared__ __align__(8) float2 s_sym[2][FB][ROW];
    __shared__ __align__(16) float4 s_out[kThreads / 32][32 * SPS / 2]; /* per warp: 32 symbols x 8 samples x 8 B */

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t i = tid; i < a.n_tables * a.n_const; i += kThreads) s_lut[i] = a.lut[i];
    const f32x2 one = pk2(taps.one.x, taps.one.y);

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);
    const u64 k0 = (u64)blockIdx.x * kThreads; /* first symbol of the tile */
    const u64 m = k0 + tid;                    /* this thread's symbol */

    float cs[SPS], sn[SPS];
    const float2* tab = chan_table(a.ch, f0);
    if (tab && (m + 1) * SPS <= a.L) {
#pragma unroll
        for (int p = 0; p < SPS; p += 2) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(tab + m * SPS + p));
            cs[p] = t.x; sn[p] = t.y; cs[p + 1] = t.z; sn[p + 1] = t.w;
        }
    } else {
#pragma unroll
        for (int p = 0; p < SPS; ++p) mg_sincosf(nco_phase(w, a.sample0 + m * SPS + p), &sn[p], &cs[p]);
    }

    /* mapped symbols of FB frames: s_sym[buf][g][i] = (i, q) of symbol k0 - HALO + i of frame fg + g.  A thread
     * stages row entry `tid` (and, the first HALO threads, entry kThreads + tid); which symbols those are, whether
     * they exist and which table they use does not depend on the frame.  The fetch of the NEXT group's bits is
     * issued before the FIR of the current one and only consumed (table lookup, shared store) after it: in the
     * first form the lookup directly behind the load stalled every warp on the global latency, then on the
     * barrier (ncu: 30 % + 27 % of the stall samples, profiles/r01_c3_signform_ncu.txt). */
    const uint32_t bps = a.bps;
    /* BPSW > 0: a symbol's BPSW bytes are one aligned word (checked by the launcher); 0: byte-wise fallback */
    constexpr bool word = BPSW > 0;
    using raw_t = typename std::conditional<BPSW == 8, unsigned long long, uint32_t>::type;
    bool in_e[2];
    uint32_t toff_e[2];
    const uint8_t* src_e[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const long long mm = (long long)k0 - HALO + tid + e * kThreads;
        in_e[e] = (e == 0 || tid < HALO) && mm >= 0 && (u64)mm < a.nsym;
        toff_e[e] = (in_e[e] && a.n_tables > 1) ? (uint32_t)((u64)mm % a.n_tables) * a.n_const : 0u;
        src_e[e] = a.bits + (in_e[e] ? (u64)mm * bps : 0);
    }
    raw_t raw[2][FB]; /* the symbols' bytes as loaded (word) or already packed (byte-wise fallback) */
    auto fetch = [&](u64 fg) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int g = 0; g < FB; ++g) {
                raw_t r = 0;
                if (in_e[e] && fg + g < f1) {
                    const uint8_t* p = src_e[e] + (fg + g) * a.nbits;
                    if (BPSW == 0) r = pack_symbol(p, bps);
                    else if (BPSW == 1) r = __ldg(p);
                    else if (BPSW == 2) r = __ldg(reinterpret_cast<const uint16_t*>(p));
                    else if (BPSW == 4) r = __ldg(reinterpret_cast<const uint32_t*>(p));
                    else r = (raw_t)__ldg(reinterpret_cast<const unsigned long long*>(p));
                }
                raw[e][g] = r;
            }
    };
    auto commit = [&](int buf, u64 fg) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (e == 1 && tid >= HALO) break;
#pragma unroll
            for (int g = 0; g < FB; ++g) {
                float2 v = make_float2(0.0f, 0.0f);
                if (in_e[e] && fg + g < f1) {
                    uint32_t idx = (uint32_t)raw[e][g];
                    if (word) { /* digital/util.rs:5-11: first byte is the MSB */
                        idx = 0;
#pragma unroll
                        for (int j = 0; j < BPSW; ++j) idx = (idx << 1) | ((uint32_t)(raw[e][g] >> (8 * j)) & 1u);
                    }
                    v = s_lut[toff_e[e] + idx];
                }
                s_sym[buf][g][tid + e * kThreads] = v;
            }
        }
    };
    __syncthreads();
    if (f0 < f1) {
        fetch(f0);
        commit(0, f0);
    }
    __syncthreads();

    int buf = 0;
    for (u64 fg = f0; fg < f1; fg += FB, buf ^= 1) {
        const bool more = fg + FB < f1;
        if (more) fetch(fg + FB); /* in flight during the FIR */
        f32x2 acc[FB][SPS];
#pragma unroll
        for (int g = 0; g < FB; ++g)
#pragma unroll
            for (int p = 0; p < SPS; ++p) acc[g][p] = 0ull; /* (+0.0f, +0.0f) */
#pragma unroll
        for (int j = 0; j < J; ++j) {
            f32x2 win[FB]; /* (i, q) of symbol m - j, one packed pair per frame */
#pragma unroll
            for (int g = 0; g < FB; ++g) win[g] = reinterpret_cast<const f32x2*>(s_sym[buf][g])[tid + HALO - j];
#pragma unroll
            for (int p = 0; p < SPS; ++p) {
                if (p + j * SPS < NT) {
                    const f32x2 hh = pk2(taps.hh[p + j * SPS].x, taps.hh[p + j * SPS].y);
#pragma unroll
                    for (int g = 0; g < FB; ++g) acc[g][p] = mac2<FMA>(acc[g][p], win[g], hh, one);
                }
            }
        }

        float4* wout = s_out[wid];
        const u64 sym_w0 = k0 + (u64)wid * 32;
#pragma unroll
        for (int g = 0; g < FB; ++g) {
            if (fg + g < f1) { /* uniform across the CTA */
#pragma unroll
                for (int pp = 0; pp < SPS; pp += 2) {
                    const float2 b0 = unpk2(acc[g][pp]), b1 = unpk2(acc[g][pp + 1]);
                    const float2 o0 = mix_iq(b0.x, b0.y, cs[pp], sn[pp]);
                    const float2 o1 = mix_iq(b1.x, b1.y, cs[pp + 1], sn[pp + 1]);
                    /* chunk c = pp/2 of row `lane` (4 chunks of 16 B per row), XOR-swizzled */
                    const int c = pp >> 1;
                    wout[lane * 4 + (c ^ ((lane >> 1) & 3))] = make_float4(o0.x, o0.y, o1.x, o1.y);
                }
                __syncwarp();
                /* coalesced write-out of this warp's 32 symbols = 256 samples = 128 float4 */
                float4* gout = reinterpret_cast<float4*>(a.tx + (fg + g) * a.L + sym_w0 * SPS);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = lane + 32 * i; /* linear 16-byte chunk in the warp tile */
                    const int row = q >> 2, c = q & 3;
                    const float4 v = wout[row * 4 + (c ^ ((row >> 1) & 3))];
                    if (sym_w0 + row < a.nsym) __stcs(gout + q, v);
                }
                __syncwarp();
            }
        }
        if (more) commit(buf ^ 1, fg + FB);
        __syncthreads(); /* s_sym[buf^1] staged, s_sym[buf] free for the group after next */
    }
}

/* ------------------------------------------------------------------ launchers */
bool tx_rect_fast_supported(uint32_t bps) { return bps == 1 || bps == 2 || bps == 4 || bps == 8; }
uint64_t tx_rect_fast_tiles(uint64_t L) { return (L + 4 * kThreads - 1) / (4 * kThreads); }
cudaError_t tx_rect_fast_launch(const TxArgs& a, cudaStream_t stream)
{
    dim3 grid((unsigned)tx_rect_fast_tiles(a.L), (unsigned)((a.F + a.frames_per_block - 1) / a.frames_per_block));
    const bool real = a.re && !a.tx, two = (a.sps & 1u) != 0;
#define MG_TXR(B)                                                                              \
    do {                                                                                       \
        if (real) {                                                                            \
            if (two) tx_rect_fast_kernel<B, true, true><<<grid, kThreads, 0, stream>>>(a);     \
            else tx_rect_fast_kernel<B, true, false><<<grid, kThreads, 0, stream>>>(a);        \
        } else {                                                                               \
            if (two) tx_rect_fast_kernel<B, false, true><<<grid, kThreads, 0, stream>>>(a);    \
            else tx_rect_fast_kernel<B, false, false><<<grid, kThreads, 0, stream>>>(a);       \
        }                                                                                      \
    } while (0)
    switch (a.bps) {
    case 1: MG_TXR(1); break;
    case 2: MG_TXR(2); break;
    case 4: MG_TXR(4); break;
    default: MG_TXR(8); break;
    }
#undef MG_TXR
    return cudaGetLastError();
}

bool tx_shaped_fast_supported(uint32_t sps, uint32_t n_taps) { return sps == 8 && n_taps == 129; }
uint64_t tx_shaped_fast_tiles(uint64_t nsym) { return (nsym + kThreads - 1) / kThreads; }
cudaError_t tx_shaped_fast_launch(const TxArgs& a, const float* h_taps, bool fma, bool rail_pairs, cudaStream_t stream)
{
    dim3 grid((unsigned)tx_shaped_fast_tiles(a.nsym), (unsigned)((a.F + a.frames_per_block - 1) / a.frames_per_block));
    /* rail_pairs: h_taps holds (h_i[k], h_q[k]) pairs and a.lut holds (+-1, +-1) -- the sign-product form, see modem_api.cu */
    const TapsParam<129> tp = rail_pairs ? make_taps_param_pairs<129>(h_taps) : make_taps_param<129>(h_taps);
    /* two frames per tap fetch: measured best (four: exact 3.65 ms against 3.24 ms at C3, fused 2.99 against 2.58) */
    const uint32_t bps = a.bps;
    const bool word = (bps == 1 || bps == 2 || bps == 4 || bps == 8) && (a.nbits % bps == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.bits) % bps) == 0);
#define MG_TXS(W)                                                                              \
    do {                                                                                       \
        if (fma) tx_shaped_fast_kernel<8, 129, true, 2, W><<<grid, kThreads, 0, stream>>>(a, tp);  \
        else tx_shaped_fast_kernel<8, 129, false, 2, W><<<grid, kThreads, 0, stream>>>(a, tp);     \
    } while (0)
    switch (word ? bps : 0u) {
    case 1: MG_TXS(1); break;
    case 2: MG_TXS(2); break;
    case 4: MG_TXS(4); break;
    case 8: MG_TXS(8); break;
    default: MG_TXS(0); break;
    }
#undef MG_TXS
    return cudaGetLastError();
}

} /* namespace mg */
