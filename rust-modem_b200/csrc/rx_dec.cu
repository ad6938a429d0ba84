/*
 * rx_dec.cu -- tuned decimating RX for ANY samples-per-symbol count (the reference's default rates give 45:
 * rates.rs:16 with src/bin/modulate.rs:44-58) and the reference's 64-tap low-pass (src/bin/demodulate.rs:82-147).
 * Replaces demodulator.rs:44-55 + two fir.rs:18-34 filters of the reference plus the decimator / slicer / error-count
 * extension, like rx_fast.cuh does for 8 samples per symbol; the generic kernel (kernels.cuh) keeps the shapes this one
 * does not take (per-frame PLL offsets, raw wire formats, an OQPSK rail offset, noise, odd frame lengths).
 *
 * At sps = 45 the FIR is 2.8 MACs per sample (64 taps x 2 rails per 45 samples): the kernel is a stream of 8 B/sample
 * from HBM and must be organised like one.
 *   phase A  the tile's samples are read two at a time (128-bit loads, eight per thread in flight before the first use; the
 *            next frame's tile is pulled towards L2 by one bulk prefetch per CTA meanwhile),
 *            their real parts multiplied by the NCO pair (cos, sin) from the context's table (ChannelView::cs_tab, read
 *            through L1/L2: frame-invariant) and staged as packed (x cos, x sin) pairs, one 128-bit shared store per
 *            sample pair.  The Q rail is staged NEGATED (x sin instead of x (-sin)): see rx_fast.cuh -- exact, and
 *            phase C takes 0 - acc.
 *   phase B  one thread = one symbol, both rails packed: 64 ordered MACs (FMUL2 + FFMA2 with (1, 1), products one step
 *            ahead of their accumulation), each fed by one 8-byte shared load; the taps are uniform operands from the
 *            constant bank.  Threads read the staged row at a stride of sps pairs: conflict-free for odd sps.
 *   phase C  slice (sign test for an axis-aligned 4-point table, else nearest point), emit, count.
 */
#include <atomic>

#include "launch.h"
#include "rx_fast.cuh" /* the tensor-memory helpers */

namespace mg {

constexpr int kDecThreads = 128; /* = symbols per tile */

/* TMEM: the frame-invariant NCO values of a thread's first 32 chunks (128 values) are parked in tensor memory once per CTA
 * (128 columns x 4 CTAs = the SM's 512 columns, all 128 lanes used) and read back with tcgen05.ld every frame, instead of
 * 8 B/sample from the NCO table through L2/L1 next to the 8 B/sample of the signal itself. */
/* (cos, sin) of phase + offset as one out-of-line routine (inlined copies of the binary64 routine stall the kernel on
 * instruction fetch: rx_fast.cuh) */
static __device__ __noinline__ float2 dec_sincos_nco(float y)
{
    float s, c;
    mg_sincosf_nco(y, &s, &c);
    return make_float2(c, s);
}

/* MODE 0 = complex samples in, 1 = TXF (the fused loopback), 2 / 3 = RAW: the real f32 / i16 wire of src/bin/demodulate.rs:29-43
 * with one PLL phase offset per frame (Demodulator::lock_phase): the frame-invariant phase(n) is what is parked in tensor
 * memory, every frame adds its offset and evaluates glibc's cosf / sinf while staging (demodulator.rs:51-54); one load per
 * sample, because the reference's odd preamble gives rows with odd strides and odd lengths. */
template <int NT, bool FMA, bool TMEM, int MODE = 0>
__global__ void __launch_bounds__(kDecThreads, 4)
    rx_dec_kernel(const __grid_constant__ RxArgs a, const __grid_constant__ TapsParam<NT> taps)
{
    constexpr bool TXF = MODE == 1;
    constexpr int RAW = MODE >= 2 ? MODE - 1 : 0; /* 1 = f32 rows, 2 = i16 rows */
    static_assert(RAW == 0 || (TMEM && !FMA), "the raw-wire variant parks its phases in tensor memory");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    f32x2* s_v = reinterpret_cast<f32x2*>(smem_raw); /* staged (x cos, x sin) pairs, index = sample - nb0 */
    __shared__ float2 s_slut[kMaxLut];
    const int tid = threadIdx.x & (kDecThreads - 1);
    for (uint32_t i = tid; i < a.n_tables * a.n_const; i += kDecThreads) s_slut[i] = a.slut[i];

    const uint32_t TS = a.sym_tile, sps = a.sps;
    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const u64 k0 = (u64)blockIdx.x * TS;
    /* first sample the tile's first symbol needs (may be negative: zero history, fir.rs:13), floored to an even index so
     * that sample pairs are the 16-byte pairs of the row */
    const long long nb = (long long)(k0 * sps + a.delay) - (long long)(NT - 1);
    const long long nb0 = nb & ~1ll;
    const uint32_t shift = (uint32_t)(nb - nb0);
    const uint32_t R = (TS - 1) * sps + NT + shift; /* samples staged */
    const uint32_t NCH = (R + 1) / 2;               /* 16-byte chunks */
    const float4* cs4 = RAW ? nullptr : reinterpret_cast<const float4*>(chan_table(a.ch, f0) + nb0);
    const float raw_w = RAW ? chan_w(a.ch, f0) : 0.0f;
    const f32x2 one = pk2(taps.one.x, taps.one.y);
    const u64 k = k0 + tid;
    const bool live = (uint32_t)tid < TS && k < a.K;
    const f32x2* mine = s_v + (size_t)tid * sps + (NT - 1) + shift; /* the decision instant of this thread's symbol */
    const bool lut4 = a.n_const == 4 && a.n_tables == 1;
    const uint32_t toff = (uint32_t)(k % a.n_tables) * a.n_const;
    const bool pair_io = a.bps == 2 && (reinterpret_cast<uintptr_t>(a.bits) & 1u) == 0 && (reinterpret_cast<uintptr_t>(a.ref_bits) & 1u) == 0 &&
                         (a.ref_stride & 1u) == 0;

    constexpr int U = 8;        /* chunks per thread and trip: all loads of a trip are in flight before the first use */
    constexpr int TRIPS_TM = 4; /* trips whose NCO values live in tensor memory: 4 x 8 chunks x 4 values = 128 columns */
    __shared__ uint32_t s_tmem;
    uint32_t taddr = 0, twarp = 0;
    /* the staged chunks that exist in the frame, as a range of chunk indices [vlo, vlo + vspan): L and nb0 are even, so a
     * 16-byte pair is inside the frame or outside; one unsigned compare per chunk */
    const uint32_t vlo = nb0 < 0 ? (uint32_t)((-nb0) >> 1) : 0u;
    /* first chunk behind the frame's end (> 0: the tile has a live symbol); a raw frame may hold an odd number of samples: its
     * last chunk then has a first sample only (vspan1 covers the chunks whose SECOND sample exists) */
    const long long vend = ((long long)a.L - nb0 + (RAW ? 1 : 0)) >> 1;
    const long long vend1 = ((long long)a.L - nb0) >> 1;
    const uint32_t vhi = vend < (long long)NCH ? (uint32_t)(vend < 0 ? 0 : vend) : NCH;
    const uint32_t vspan = vhi > vlo ? vhi - vlo : 0u;
    auto chunk_ok = [&](uint32_t c) { return c - vlo < vspan; };
    const uint32_t vhi1 = vend1 < (long long)NCH ? (uint32_t)(vend1 < 0 ? 0 : vend1) : NCH;
    const uint32_t vspan1 = vhi1 > vlo ? vhi1 - vlo : 0u;
    auto chunk_ok1 = [&](uint32_t c) { return c - vlo < vspan1; };
    /* ---- TXF: the fused loopback for any samples-per-symbol count (the reference's default rates).  Phase A does not load
     * the tile's TX samples, it MAKES them from the frame's bits (a.ref_bits, two bytes per symbol) with the arithmetic of
     * the rectangular-hold TX kernel -- data.rs:66-79 hold, digital/qpsk.rs:23-35 as a 4-entry table (a.tx_iq),
     * modulator.rs:37-48 mix, every operation rounded on its own --, stores them to a.tx_out once (a tile owns the samples
     * from the end of the previous tile's staged range to the end of its own; the first tile owns the frame's head, the
     * last one its tail: the launcher checks that the staged ranges leave no gap, sps <= NT) and stages their real parts
     * like the unfused kernel.  The symbols a tile touches (at most 2 * kDecThreads: its own 128 plus (NT + 2) / sps + 2 of
     * halo) reach the threads as (i, q) pairs through a small shared table, written when the previous frame's FIR is done
     * from bit bytes fetched before it started.  Same NCO values on both sides (no phase offset: the launcher checks). */
    __shared__ f32x2 s_sq[TXF ? 2 * kDecThreads : 1];
    const long long first_n = nb0 < 0 ? 0 : nb0;                 /* first staged sample that exists */
    const uint32_t ksym0 = TXF ? (uint32_t)((u64)first_n / sps) : 0u; /* table slot 0 */
    const uint32_t magic = TXF ? 0xFFFFFFFFu / sps + 1u : 0u;    /* ceil(2^32 / sps): floor(x / sps) = umulhi(x, magic) for x < 2^32 / sps */
    const long long loc0 = nb0 - (long long)ksym0 * sps;         /* sample nb0 relative to the first sample of symbol ksym0 (negative in the first tile) */
    uint32_t olo = 0, ospan = 0; /* the chunks whose TX samples this tile stores: [olo, olo + ospan), inside [vlo, vhi) */
    if (TXF) {
        long long own_lo = 0;
        if (k0 > 0) { /* the previous tile's staged range ends here */
            const long long pnb = (long long)((k0 - TS) * sps + a.delay) - (long long)(NT - 1);
            const long long pnb0 = pnb & ~1ll;
            const uint32_t pR = (TS - 1) * sps + NT + (uint32_t)(pnb - pnb0);
            own_lo = pnb0 + 2 * (long long)((pR + 1) / 2);
        }
        const long long lo_c = (own_lo - nb0) >> 1; /* own_lo is even and >= max(nb0, 0) */
        olo = lo_c < (long long)vlo ? vlo : (uint32_t)lo_c;
        ospan = vhi > olo ? vhi - olo : 0u; /* a tile stores up to the end of its staged range; the last one reaches the frame's end */
    }
    const uint32_t loc0u = (uint32_t)loc0; /* modulo 2^32: only sums that are >= 0 are used */
    const u64 nsym = TXF ? a.ref_stride / 2 : 0; /* symbols per frame row (bps = 2) */
    const bool store_tx = TXF && a.tx_out != nullptr;
    uint32_t nw0 = 0, nw1 = 0; /* the next frame's bit bytes of this thread's two table slots */
    auto fetch_syms = [&](u64 f) {
        const uint8_t* row = a.ref_bits + f * a.ref_stride;
        const u64 s0 = (u64)ksym0 + tid, s1 = s0 + kDecThreads;
        nw0 = s0 < nsym ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(row + 2 * s0)) : 0u;
        nw1 = s1 < nsym ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(row + 2 * s1)) : 0u;
    };
    auto park_syms = [&]() { /* symbol index = 2 b0 + b1, first byte = MSB (digital/util.rs:5-11); only bit 0 of a byte counts */
        const float2 p0 = a.tx_iq[((nw0 & 1u) << 1) | ((nw0 >> 8) & 1u)], p1 = a.tx_iq[((nw1 & 1u) << 1) | ((nw1 >> 8) & 1u)];
        s_sq[TXF ? tid : 0] = pk2(p0.x, p0.y);
        s_sq[TXF ? tid + kDecThreads : 0] = pk2(p1.x, p1.y);
    };
    if (TXF && f0 < f1) {
        fetch_syms(f0);
        park_syms();
    }
    if (TMEM) {
        taddr = tmem_alloc<128>(&s_tmem);
        twarp = tmem_warp_addr(taddr);
#pragma unroll
        for (int b = 0; b < TRIPS_TM; ++b) {
            float park[32];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t c = tid + (b * U + u) * kDecThreads;
                float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (chunk_ok(c)) {
                    if (RAW) { /* (phase 0, phase 1, -, -): carrier.rs:17-19 */
                        const u64 n = (u64)(nb0 + 2 * (long long)c);
                        t.x = nco_phase(raw_w, a.sample0 + n);
                        t.y = nco_phase(raw_w, a.sample0 + n + 1);
                    } else {
                        t = __ldg(cs4 + c);
                    }
                }
                park[4 * u] = t.x; park[4 * u + 1] = t.y; park[4 * u + 2] = t.z; park[4 * u + 3] = t.w;
            }
            tmem_st32(twarp + 32 * b, park);
        }
    }

    uint32_t err = 0, cmp = 0;
    for (u64 f = f0; f < f1; ++f) {
        __syncthreads(); /* previous frame's FIR finished; s_slut visible */
        const float4* src = RAW ? nullptr : reinterpret_cast<const float4*>(a.rx + f * a.L + nb0);
        /* RAW: element index of the frame's sample nb0 in the wire rows, this frame's PLL offset, and the sample loader */
        const long long raw_e0 = RAW ? (long long)(f * a.raw_stride + a.raw_skip) + nb0 : 0;
        const float raw_po = RAW ? chan_po(a.ch, f) : 0.0f;
        auto raw_load = [&](uint32_t c) -> float4 { /* (x0, -, x1, -) like the real parts of a complex pair */
            float x0 = 0.0f, x1 = 0.0f;
            const long long e = raw_e0 + 2 * (long long)c;
            if (RAW == 1) {
                const float* w = reinterpret_cast<const float*>(a.raw);
                if (chunk_ok(c)) x0 = __ldg(w + e);
                if (chunk_ok1(c)) x1 = __ldg(w + e + 1);
            } else if (RAW == 2) { /* `x as f32` of an i16 sample (demodulate.rs:29) */
                const short* w = reinterpret_cast<const short*>(a.raw);
                if (chunk_ok(c)) x0 = (float)__ldg(w + e);
                if (chunk_ok1(c)) x1 = (float)__ldg(w + e + 1);
            }
            return make_float4(x0, 0.0f, x1, 0.0f);
        };
        /* RAW: (cos, sin) of the chunk's two samples from their parked phases: phase = carrier.next() + pll.phase_offset
         * (demodulator.rs:51), evaluated for samples that exist only */
        auto raw_cs = [&](uint32_t c, float p0, float p1) -> float4 {
            float4 cs = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (chunk_ok(c)) { const float2 t = dec_sincos_nco(__fadd_rn(p0, raw_po)); cs.x = t.x; cs.y = t.y; }
            if (chunk_ok1(c)) { const float2 t = dec_sincos_nco(__fadd_rn(p1, raw_po)); cs.z = t.x; cs.w = t.y; }
            return cs;
        };
        /* pull the NEXT frame's tile towards L2 while this one is staged and filtered: one bulk (TMA) prefetch per CTA */
        if (!TXF && !RAW && tid == 0 && f + 1 < f1) {
            const long long lo = nb0 < 0 ? 0 : nb0, hi = min((long long)a.L, nb0 + 2 * (long long)NCH);
            if (hi > lo) {
                const char* nxt = reinterpret_cast<const char*>(a.rx + (f + 1) * a.L + lo);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"((int)(hi - lo) * 8) : "memory");
            }
        }
        auto stage = [&](uint32_t c, const float4& x, const float4& cs) { /* demodulator.rs:53-54, Q rail negated (see the header) */
            reinterpret_cast<ulonglong2*>(s_v)[c] = make_ulonglong2(mul2(pk2(x.x, x.x), pk2(cs.x, cs.y)), mul2(pk2(x.z, x.z), pk2(cs.z, cs.w)));
        };
        float4* txrow = TXF ? reinterpret_cast<float4*>(a.tx_out + f * a.L + nb0) : nullptr;
        /* TXF: the chunk's two TX samples from the symbol table and the NCO pair (cs = cos, sin of samples 2c, 2c + 1; zero
         * outside the frame, where nothing is made or stored) */
        auto make_tx = [&](uint32_t c, const float4& cs) -> float4 {
            /* no branch for chunks outside the frame: their NCO pairs are zero (the staged values become +-0, which no FIR sum
             * notices), their table slot is clamped into the table (every slot holds a finite pair) and they are never stored */
            const uint32_t l0 = loc0u + 2u * c; /* >= 0 for a chunk inside the frame */
            constexpr uint32_t SMAX = 2 * kDecThreads - 1;
            const float2 q0 = unpk2(s_sq[TXF ? min(__umulhi(l0, magic), SMAX) : 0]), q1 = unpk2(s_sq[TXF ? min(__umulhi(l0 + 1u, magic), SMAX) : 0]);
            const f32x2 pm = pk2(-1.0f, 1.0f);
            /* modulator.rs:37-43 on packed pairs: (i*c, i*s) and (q*s, q*c), then (i*c - q*s, i*s + q*c) by one fma with (-1, +1) */
            const f32x2 x0 = fma2(mul2(pk2(q0.y, q0.y), pk2(cs.y, cs.x)), pm, mul2(pk2(q0.x, q0.x), pk2(cs.x, cs.y)));
            const f32x2 x1 = fma2(mul2(pk2(q1.y, q1.y), pk2(cs.w, cs.z)), pm, mul2(pk2(q1.x, q1.x), pk2(cs.z, cs.w)));
            const float4 x = make_float4(unpk2(x0).x, unpk2(x0).y, unpk2(x1).x, unpk2(x1).y);
            if (store_tx && c - olo < ospan) __stcs(txrow + c, x);
            return x;
        };
        uint32_t c0 = tid;
        if (TMEM) {
#pragma unroll
            for (int b = 0; b < TRIPS_TM; ++b, c0 += U * kDecThreads) {
                if (c0 - tid >= NCH) break; /* uniform: the tile has no more chunks */
                /* FULL: every chunk of the trip is a staged chunk (uniform), no per-chunk range test */
                auto trip = [&](auto full_c) {
                    constexpr bool FULL = decltype(full_c)::value;
                    float4 x[U];
                    if (!TXF) {
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const uint32_t c = c0 + u * kDecThreads;
                            if (RAW) x[u] = raw_load(c);
                            else x[u] = chunk_ok(c) ? __ldg(src + c) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        }
                    }
                    float parked[32];
                    tmem_ld32(twarp + 32 * b, parked);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t c = c0 + u * kDecThreads;
                        const float4 cs = RAW ? raw_cs(c, parked[4 * u], parked[4 * u + 1])
                                              : make_float4(parked[4 * u], parked[4 * u + 1], parked[4 * u + 2], parked[4 * u + 3]);
                        if (FULL || c < NCH) stage(c, TXF ? make_tx(c, cs) : x[u], cs);
                    }
                };
                if (c0 - tid + U * kDecThreads <= NCH) trip(std::true_type{});
                else trip(std::false_type{});
            }
        }
        for (; c0 < NCH; c0 += U * kDecThreads) { /* chunks beyond the parked ones (large sps), or all of them without TMEM */
            float4 x[U], cs[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t c = c0 + u * kDecThreads;
                const bool ok = chunk_ok(c);
                if (RAW) { /* chunks beyond the parked ones (sps > 64): their phases are evaluated here */
                    x[u] = raw_load(c);
                    const u64 n = (u64)(nb0 + 2 * (long long)c);
                    cs[u] = ok ? raw_cs(c, nco_phase(raw_w, a.sample0 + n), nco_phase(raw_w, a.sample0 + n + 1)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    continue;
                }
                x[u] = (ok && !TXF) ? __ldg(src + c) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                cs[u] = ok ? __ldg(cs4 + c) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t c = c0 + u * kDecThreads;
                if (c < NCH) stage(c, TXF ? make_tx(c, cs[u]) : x[u], cs[u]);
            }
        }
        __syncthreads();
        if (TXF && f + 1 < f1) fetch_syms(f + 1); /* the next frame's bit bytes fly during the FIR */
        if (live) {
            /* fir.rs:21-24 on both rails at once: acc = fl(acc + fl(v * h)), taps 0..NT-1 in order */
            f32x2 acc = 0ull, pend = 0ull;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const f32x2 v = mine[-i];
                const f32x2 hh = pk2(taps.hh[i].x, taps.hh[i].y);
                if (FMA) {
                    acc = fma2(v, hh, acc);
                } else {
                    const f32x2 prod = mul2(v, hh);
                    if (i > 0) acc = fma2(acc, one, pend);
                    pend = prod;
                }
            }
            if (!FMA) acc = fma2(acc, one, pend);
            const float2 t = unpk2(acc);
            const float I = __fmul_rn(a.rx_gain, t.x), Q = __fmul_rn(a.rx_gain, __fsub_rn(0.0f, t.y));
            uint32_t s;
            if (a.sign_slice && fabsf(I) >= a.ss_lo && fabsf(I) <= a.ss_hi && fabsf(Q) >= a.ss_lo && fabsf(Q) <= a.ss_hi)
                s = ((~__float_as_uint(I) >> 31) << 1) | (~__float_as_uint(Q) >> 31);
            else
                s = lut4 ? slice_point(s_slut, 4, I, Q) : slice_point(s_slut + toff, a.n_const, I, Q);
            if (pair_io) { /* bps = 2: the two bit bytes of a symbol as one 16-bit word, first byte = MSB (digital/util.rs:5-11) */
                const u64 o = f * a.K + k;
                const uint32_t bw = (s >> 1) | ((s & 1u) << 8);
                if (a.sym) a.sym[o] = (uint8_t)s;
                if (a.soft) a.soft[o] = make_float2(I, Q);
                if (a.bits) *reinterpret_cast<unsigned short*>(a.bits + 2 * o) = (unsigned short)bw;
                if (a.ref_bits) /* only bit 0 of a reference byte counts, as in pack_symbol */
                    err += __popc((bw ^ (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(a.ref_bits + f * a.ref_stride + 2 * k))) & 0x0101u);
            } else {
                err += emit_symbol(a, f, k, s, I, Q);
            }
            cmp += a.ref_bits ? a.bps : 0u;
        }
        if (TXF && f + 1 < f1) park_syms(); /* read again only behind the barrier at the top of the next frame */
    }
    block_count(a, err, cmp);
    if (TMEM) tmem_free<128>(taddr);
}

bool rx_dec_supported(uint32_t n_taps, uint32_t sps) { return n_taps == 64 && sps >= 2 && sps <= 180; }
uint32_t rx_dec_tile_symbols(uint32_t sps)
{
    (void)sps;
    return kDecThreads; /* (127 sps + 66) * 8 B of shared memory: 46 KB at sps 45, 183 KB at sps 180 */
}
template <int NT, bool FMA, bool TMEM, int MODE = 0>
static cudaError_t rx_dec_launch_t(const RxArgs& a, const float* h_taps, cudaStream_t stream)
{
    const size_t smem = ((size_t)(a.sym_tile - 1) * a.sps + NT + 2) * sizeof(f32x2);
    auto kern = rx_dec_kernel<NT, FMA, TMEM, MODE>;
    static std::atomic<size_t> configured[kMaxDevices]; /* per device: the attribute is per device */
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices || configured[dev].load(std::memory_order_acquire) != smem + 1) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < kMaxDevices) configured[dev].store(smem + 1, std::memory_order_release);
    }
    dim3 grid((unsigned)((a.K + a.sym_tile - 1) / a.sym_tile), (unsigned)((a.F + a.frames_per_block - 1) / a.frames_per_block));
    kern<<<grid, kDecThreads, smem, stream>>>(a, make_taps_param<NT>(h_taps));
    return cudaGetLastError();
}
cudaError_t rx_dec_launch(const RxArgs& a, const float* h_taps, bool fma, bool tmem, cudaStream_t stream)
{
    if (tmem) return fma ? rx_dec_launch_t<64, true, true>(a, h_taps, stream) : rx_dec_launch_t<64, false, true>(a, h_taps, stream);
    return fma ? rx_dec_launch_t<64, true, false>(a, h_taps, stream) : rx_dec_launch_t<64, false, false>(a, h_taps, stream);
}

/* ---- the fused loopback at any samples-per-symbol count (rx_dec_kernel<..., TXF>): geometry test and launcher.
 * a.L, a.K, a.sps, a.delay and a.sym_tile must be set.  The tiles' staged ranges must cover [0, L) without a gap: the
 * first tile starts at or before sample 0 (delay <= NT - 1), consecutive tiles overlap (sps <= NT) and the last one
 * reaches the frame's end. */
bool loop_fused_dec_supported(const RxArgs& a)
{
    constexpr uint32_t NT = 64;
    const uint32_t sps = a.sps, TS = kDecThreads;
    if (!rx_dec_supported(NT, sps) || sps > NT || a.delay > NT - 1 || a.K == 0 || (a.L & 1) || a.L >= (1ull << 32) / sps) return false;
    const u64 tiles = (a.K + TS - 1) / TS;
    const long long nb = (long long)((tiles - 1) * TS * sps + a.delay) - (long long)(NT - 1), nb0 = nb & ~1ll;
    const long long R = (long long)(TS - 1) * sps + NT + (nb - nb0);
    return nb0 + 2 * ((R + 1) / 2) >= (long long)a.L;
}
cudaError_t loop_fused_dec_launch(const RxArgs& a, const float* h_taps, bool tmem, cudaStream_t stream)
{
    if (tmem) return rx_dec_launch_t<64, false, true, 1>(a, h_taps, stream);
    return rx_dec_launch_t<64, false, false, 1>(a, h_taps, stream);
}
/* ---- the demodulate binary's wire at any samples-per-symbol count (the reference's own rates: 45): real f32 (fmt 1) / i16
 * (fmt 2) rows, one PLL offset per frame */
bool rx_dec_raw_supported(uint32_t n_taps, uint32_t sps, uint32_t fmt) { return rx_dec_supported(n_taps, sps) && (fmt == 1 || fmt == 2); }
cudaError_t rx_dec_raw_launch(const RxArgs& a, const float* h_taps, cudaStream_t stream)
{
    if (a.rx_fmt == 1) return rx_dec_launch_t<64, false, true, 2>(a, h_taps, stream);
    return rx_dec_launch_t<64, false, true, 3>(a, h_taps, stream);
}

} /* namespace mg */
