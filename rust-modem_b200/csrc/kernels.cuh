/*
 * kernels.cuh -- the generic (any sps / tap count / scheme) kernels and the AWGN kernel.
 * The compile-time specialised hot kernels live in tx_fast.cu and rx_fast.cuh.
 */
#pragma once

#include "common.cuh"

namespace mg {

/*
 * Rectangular-hold TX (the reference's own pulse).  One thread produces VEC consecutive
 * samples per step (VEC=2 -> one 128-bit store), U steps per tile, then loops over frames.
 */
template <int VEC>
__global__ void __launch_bounds__(kThreads) tx_rect_kernel(const __grid_constant__ TxArgs a)
{
    constexpr int U = 2;
    __shared__ float2 s_lut[kMaxLut];
    for (uint32_t i = threadIdx.x; i < a.n_tables * a.n_const; i += kThreads) s_lut[i] = a.lut[i];
    __syncthreads();

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);

    u64 n0[U];
    uint32_t ki[U][VEC], kq[U][VEC], toff[U][VEC];
    bool qv[U][VEC];
    float cs[U][VEC], sn[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        n0[u] = (((u64)blockIdx.x * U + u) * kThreads + threadIdx.x) * VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            u64 n = n0[u] + v;
            ki[u][v] = (uint32_t)(n / a.sps);
            toff[u][v] = (ki[u][v] % a.n_tables) * a.n_const;
            qv[u][v] = n >= a.q_offset;
            kq[u][v] = qv[u][v] ? (uint32_t)((n - a.q_offset) / a.sps) : 0u;
            mg_sincosf(nco_phase(w, a.sample0 + n), &sn[u][v], &cs[u][v]);
        }
    }

    for (u64 f = f0; f < f1; ++f) {
        const uint8_t* fb = a.bits + f * a.nbits;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (n0[u] >= a.L) continue;
            float2 bb[VEC], out[VEC];
            uint32_t idx = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (a.q_offset == 0) {
                    if (v == 0 || ki[u][v] != ki[u][v - 1]) idx = pack_symbol(fb + (u64)ki[u][v] * a.bps, a.bps);
                } else {
                    uint32_t b0 = __ldg(fb + (u64)ki[u][v] * 2) & 1u;
                    uint32_t b1 = qv[u][v] ? (__ldg(fb + (u64)kq[u][v] * 2 + 1) & 1u) : 0u;
                    idx = (b0 << 1) | b1;
                }
                bb[v] = s_lut[toff[u][v] + idx];
                out[v] = mix_iq(bb[v].x, bb[v].y, cs[u][v], sn[u][v]);
            }
            const u64 o = f * a.L + n0[u];
            if (a.re) {
                float* r = a.re + f * a.re_stride + a.re_offset + n0[u];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (n0[u] + v < a.L) __stcs(r + v, out[v].x);
            }
            if (VEC == 2) {
                if (a.tx) __stcs(reinterpret_cast<float4*>(a.tx + o), make_float4(out[0].x, out[0].y, out[VEC - 1].x, out[VEC - 1].y));
                if (a.iq) __stcs(reinterpret_cast<float4*>(a.iq + o), make_float4(bb[0].x, bb[0].y, bb[VEC - 1].x, bb[VEC - 1].y));
            } else {
                if (a.tx) __stcs(a.tx + o, out[0]);
                if (a.iq) __stcs(a.iq + o, bb[0]);
            }
        }
    }
}

/*
 * Generic pulse-shaped TX (any sps, any tap count): polyphase evaluation of the FIR over
 * the zero-stuffed symbol train.  The skipped terms are exact `s + 0.0*c` no-ops in the
 * reference's fold (fir.rs:21-24), so visiting only taps k == n (mod sps) in ascending k
 * reproduces the same binary32 sum.
 * dynamic smem: float s_taps[n_taps]; float s_si[TS+H]; float s_sq[TS+H]; float2 s_cs[TS*sps]
 */
template <bool FMA>
__global__ void __launch_bounds__(kThreads) tx_shaped_generic_kernel(const __grid_constant__ TxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float2 s_lut[kMaxLut];
    const uint32_t TS = a.sym_tile, sps = a.sps, N = a.n_taps;
    const uint32_t J = (N + sps - 1) / sps, H = J + 1;
    float2* s_cs = reinterpret_cast<float2*>(smem_raw);
    float* s_taps = reinterpret_cast<float*>(s_cs + (size_t)TS * sps);
    float* s_si = s_taps + N;
    float* s_sq = s_si + (TS + H);

    for (uint32_t i = threadIdx.x; i < a.n_tables * a.n_const; i += kThreads) s_lut[i] = a.lut[i];
    for (uint32_t i = threadIdx.x; i < N; i += kThreads) s_taps[i] = a.taps[i];

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);
    const u64 k0 = (u64)blockIdx.x * TS;              /* first symbol of the tile */
    const u64 nb = k0 * sps;                          /* first sample of the tile */
    const u64 ne = min(a.L, nb + (u64)TS * sps);
    for (u64 n = nb + threadIdx.x; n < ne; n += kThreads) {
        float s, c;
        mg_sincosf(nco_phase(w, a.sample0 + n), &s, &c);
        s_cs[n - nb] = make_float2(c, s);
    }
    const long long mbase = (long long)k0 - (long long)H; /* symbol index of s_si[0] */

    for (u64 f = f0; f < f1; ++f) {
        const uint8_t* fb = a.bits + f * a.nbits;
        __syncthreads(); /* previous frame's readers done; LUT/taps visible on first pass */
        for (uint32_t i = threadIdx.x; i < TS + H; i += kThreads) {
            long long m = mbase + i;
            float vi = 0.0f, vq = 0.0f;
            if (m >= 0 && (u64)m < a.nsym) {
                const float2* lut = s_lut + ((u64)m % a.n_tables) * a.n_const;
                if (a.q_offset == 0) {
                    float2 p = lut[sym_index_plain(fb, (u64)m, a.bps)];
                    vi = p.x;
                    vq = p.y;
                } else {
                    uint32_t b0 = __ldg(fb + (u64)m * 2) & 1u, b1 = __ldg(fb + (u64)m * 2 + 1) & 1u;
                    uint32_t b1p = m > 0 ? (__ldg(fb + (u64)(m - 1) * 2 + 1) & 1u) : 0u;
                    vi = lut[(b0 << 1) | b1p].x; /* I impulse at m*sps: cur = [b0(m), b1(m-1)] */
                    vq = lut[(b0 << 1) | b1].y;  /* Q impulse at m*sps+q_offset: cur = [b0(m), b1(m)] */
                }
            }
            s_si[i] = vi;
            s_sq[i] = vq;
        }
        __syncthreads();
        for (u64 n = nb + threadIdx.x; n < ne; n += kThreads) {
            float ai = 0.0f, aq = 0.0f;
            {
                uint32_t p = (uint32_t)(n % sps);
                long long m = (long long)(n / sps) - mbase;
                for (uint32_t t = p; t < N; t += sps, --m) ai = mac<FMA>(ai, s_si[m], s_taps[t]);
            }
            if (n >= a.q_offset) {
                u64 nq = n - a.q_offset;
                uint32_t p = (uint32_t)(nq % sps);
                long long m = (long long)(nq / sps) - mbase;
                for (uint32_t t = p; t < N; t += sps, --m) aq = mac<FMA>(aq, s_sq[m], s_taps[t]);
            }
            float2 cs = s_cs[n - nb];
            const u64 o = f * a.L + n;
            if (a.iq) __stcs(a.iq + o, make_float2(ai, aq));
            if (a.tx || a.re) {
                const float2 m = mix_iq(ai, aq, cs.x, cs.y);
                if (a.tx) __stcs(a.tx + o, m);
                if (a.re) __stcs(a.re + f * a.re_stride + a.re_offset + n, m.x);
            }
        }
    }
}

/* NCO table: row c = pad_lo zeros, then (cos t, sin t) for n = 0..len-1, t = mod_trig(w_c * (sample0 + n) as f32)
 * (+ po_c if with_po), then pad_hi zeros.  carrier.rs:17-26 + util.rs:3-6 evaluated once per (channel, sample
 * index) instead of per frame; the zero margins stand for "no sample there" (zero FIR history, fir.rs:13). */
__global__ void __launch_bounds__(kThreads)
    carrier_table_kernel(float2* out, u64 len, u64 pad_lo, u64 pad_hi, u64 ch0, u64 nch, ChannelView ch, u64 sample0, int with_po)
{
    const u64 row = pad_lo + len + pad_hi, total = nch * row;
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 c = g / row, j = g % row;
        float2 v = make_float2(0.0f, 0.0f);
        if (j >= pad_lo && j < pad_lo + len) {
            const u64 n = j - pad_lo;
            const float w = ch.w ? __ldg(ch.w + ch0 + c) : ch.w0;
            const float po = with_po ? (ch.po ? __ldg(ch.po + ch0 + c) : ch.po0) : 0.0f;
            float sn, cs;
            mg_sincosf_nco(__fadd_rn(nco_phase(w, sample0 + n), po), &sn, &cs);
            v = make_float2(cs, sn);
        }
        out[g] = v;
    }
}

/* ================================================================== AWGN ========== */
/* In place: buf[f][n] += sigma * (a + j b).  One thread = one quad of samples = two Philox blocks (real, imaginary rail). */
__global__ void __launch_bounds__(kThreads) awgn_kernel(float2* buf, u64 F, u64 L, Noise nz)
{
    const u64 quads_per_frame = (L + 3) / 4;
    const u64 total = F * quads_per_frame;
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 f = g / quads_per_frame, quad = g % quads_per_frame;
        const u64 gf = nz.frame0 + f;
        uint32_t ra[4], rb[4];
        noise_quad(nz, gf, quad, 0, ra);
        noise_quad(nz, gf, quad, 1, rb);
        float a[4], b[4];
        box_muller(ra[0], ra[1], &a[0], &a[1]);
        box_muller(ra[2], ra[3], &a[2], &a[3]);
        box_muller(rb[0], rb[1], &b[0], &b[1]);
        box_muller(rb[2], rb[3], &b[2], &b[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const u64 n = quad * 4 + e;
            if (n >= L) break;
            float2* p = buf + f * L + n;
            float2 v = *p;
            v.x = __fadd_rn(v.x, __fmul_rn(nz.sigma, a[e]));
            v.y = __fadd_rn(v.y, __fmul_rn(nz.sigma, b[e]));
            *p = v;
        }
    }
}

/* Payload bits from Philox (extension, oracle/modem_oracle.h "random payload bits"): one thread = one block = 128 bits */
__global__ void __launch_bounds__(kThreads) random_bits_kernel(uint8_t* bits, u64 F, u64 nbits, u64 seed, u64 frame0)
{
    const u64 blocks_per_frame = (nbits + 127) / 128;
    const u64 total = F * blocks_per_frame;
    const bool vec = (nbits % 16 == 0) && ((reinterpret_cast<uintptr_t>(bits) & 15u) == 0);
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 f = g / blocks_per_frame, b = g % blocks_per_frame;
        const u64 gf = frame0 + f;
        uint32_t r[4];
        philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), (uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)seed,
                      (uint32_t)(seed >> 32) ^ 0x62697473u, r);
        uint8_t* row = bits + f * nbits + b * 128;
        const u64 left = nbits - b * 128;
        if (vec && left >= 128) {
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int h = 0; h < 2; ++h) { /* 16 bits -> 16 bytes */
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t nib = (r[w] >> (16 * h + 4 * k)) & 0xfu;
                        o[k] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
                    }
                    reinterpret_cast<uint4*>(row)[2 * w + h] = make_uint4(o[0], o[1], o[2], o[3]);
                }
        } else {
            for (u64 j = 0; j < left && j < 128; ++j) row[j] = (uint8_t)((r[j / 32] >> (j % 32)) & 1u);
        }
    }
}

/* Packed payloads (extension, modem_gpu_loopback_packed; oracle/modem_oracle.h "packed payload"): rows of
 * ceil(nbits/8) bytes, bit j of a frame = bit 7 - j%8 of byte j/8 (first bit = most significant, the order in which
 * digital/util.rs:5-11 packs a symbol), pad bits of a row's last byte zero.  One thread = one packed byte = one 8-byte
 * word of bit bytes (the layout data.rs:35-40 consumes). */
__global__ void __launch_bounds__(kThreads) unpack_bits_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ bits, u64 F, u64 nbits)
{
    const u64 pb = (nbits + 7) / 8, total = F * pb;
    const bool vec = (nbits % 8 == 0) && ((reinterpret_cast<uintptr_t>(bits) & 7u) == 0);
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 f = g / pb, j = g % pb;
        /* the byte replicated eight times; word byte i keeps bit 7-i of it, then any non-zero byte becomes 1 */
        unsigned long long w = ((unsigned long long)packed[g] * 0x0101010101010101ull) & 0x0102040810204080ull;
        w = ((w + 0x7f7f7f7f7f7f7f7full) >> 7) & 0x0101010101010101ull;
        uint8_t* dst = bits + f * nbits + 8 * j;
        if (vec) {
            *reinterpret_cast<unsigned long long*>(dst) = w;
        } else {
            const u64 left = nbits - 8 * j;
            for (u64 i = 0; i < 8 && i < left; ++i) dst[i] = (uint8_t)(w >> (8 * i));
        }
    }
}
__global__ void __launch_bounds__(kThreads) pack_bits_kernel(const uint8_t* __restrict__ bits, uint8_t* __restrict__ packed, u64 F, u64 nbits)
{
    const u64 pb = (nbits + 7) / 8, total = F * pb;
    const bool vec = (nbits % 8 == 0) && ((reinterpret_cast<uintptr_t>(bits) & 7u) == 0);
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 f = g / pb, j = g % pb;
        const uint8_t* src = bits + f * nbits + 8 * j;
        unsigned long long w = 0;
        if (vec) {
            w = __ldg(reinterpret_cast<const unsigned long long*>(src));
        } else {
            const u64 left = nbits - 8 * j;
            for (u64 i = 0; i < 8 && i < left; ++i) w |= (unsigned long long)src[i] << (8 * i);
        }
        /* bit 0 of word byte i lands on bit 63-i of the product: 8i + 9(7-i) = 63 - i, no two terms share a position */
        packed[g] = (uint8_t)(((w & 0x0101010101010101ull) * 0x8040201008040201ull) >> 56);
    }
}

/*
 * Generic decimating RX (any sps / tap count / q_offset): one thread per symbol.
 * dynamic smem: float2 s_cs[R]; float s_vi[RP]; float s_vq[RP]; float s_taps[N]
 *   with R = (TS-1)*sps + N + q_offset samples per tile and RP = R + R/32 + 1: element j of a rail lives at
 *   j + j/32.  The threads of a warp read elements sps apart; for sps = 4, 8, 16 the unskewed layout puts them on
 *   32/sps banks (8-way conflicts at sps 8 -- measured 4.5 ms where the skewed layout needs 2); the one-word skew
 *   per 32 elements spreads any power-of-two stride over all banks and costs nothing for odd sps.
 */
__host__ __device__ __forceinline__ uint32_t rx_generic_skew(uint32_t j) { return j + (j >> 5); }

template <bool FMA>
__global__ void __launch_bounds__(kThreads) rx_generic_kernel(const __grid_constant__ RxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float2 s_slut[kMaxLut];
    const uint32_t TS = a.sym_tile, sps = a.sps, N = a.n_taps;
    const uint32_t R = (TS - 1) * sps + N + a.q_offset;
    const uint32_t RP = rx_generic_skew(R) + 1;
    float2* s_cs = reinterpret_cast<float2*>(smem_raw);
    float* s_vi = reinterpret_cast<float*>(s_cs + R);
    float* s_vq = s_vi + RP;
    float* s_taps = s_vq + RP;

    for (uint32_t i = threadIdx.x; i < a.n_tables * a.n_const; i += kThreads) s_slut[i] = a.slut[i];
    for (uint32_t i = threadIdx.x; i < N; i += kThreads) s_taps[i] = a.taps[i];

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);
    const bool pfp = a.ch.po_frame != nullptr; /* every frame has its own PLL offset */
    const float po = pfp ? 0.0f : chan_po(a.ch, f0);
    const u64 k0 = (u64)blockIdx.x * TS;
    /* sample index of s_v*[0]; may be negative (zero history, fir.rs:13) */
    const long long nb = (long long)(k0 * sps + a.delay) - (long long)(N - 1);
    for (uint32_t j = threadIdx.x; j < R; j += kThreads) {
        long long n = nb + j;
        float s = 0.0f, c = 0.0f;
        if (n >= 0 && (u64)n < a.L) {
            const float ph = nco_phase(w, a.sample0 + (u64)n);
            if (pfp) c = ph; /* the frame-invariant part; the offset is added per frame below */
            else mg_sincosf_nco(__fadd_rn(ph, po), &s, &c);
        }
        s_cs[j] = make_float2(c, s);
    }

    uint32_t err = 0, cmp = 0;
    for (u64 f = f0; f < f1; ++f) {
        const float pof = pfp ? chan_po(a.ch, f) : 0.0f;
        __syncthreads();
        /* staging, four samples per thread and trip with the global loads issued first: a rolled one-sample loop
         * exposes the full DRAM latency once per sample (measured: 2.7 ms at the reference's sps 45) */
        for (uint32_t j0 = threadIdx.x; j0 < R; j0 += 4 * kThreads) {
            float x[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + u * kThreads;
                const long long n = nb + j;
                ok[u] = j < R && n >= 0 && (u64)n < a.L;
                x[u] = ok[u] ? rx_sample(a, f, (u64)n) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + u * kThreads;
                if (j >= R) break;
                float2 v = make_float2(0.0f, 0.0f);
                if (ok[u]) {
                    float2 cs = s_cs[j];
                    if (pfp) mg_sincosf_nco(__fadd_rn(cs.x, pof), &cs.y, &cs.x);
                    float xv = x[u];
                    if (a.nz.sigma != 0.0f) xv = __fadd_rn(xv, __fmul_rn(a.nz.sigma, noise_re(a.nz, a.nz.frame0 + f, (u64)(nb + j))));
                    v = make_float2(__fmul_rn(xv, cs.x), __fmul_rn(xv, -cs.y)); /* demodulator.rs:53-54 */
                }
                s_vi[rx_generic_skew(j)] = v.x;
                s_vq[rx_generic_skew(j)] = v.y;
            }
        }
        __syncthreads();
        /* one thread per (symbol, rail): even lanes fold the I rail, odd lanes the Q rail of the same symbol (each
         * fold is an ordered chain of N dependent MACs, fir.rs:21-24 -- two chains per symbol run side by side),
         * then the pair meets through one shuffle and the even lane slices */
        for (uint32_t t2 = threadIdx.x; t2 < ((2 * TS + 31) & ~31u); t2 += kThreads) {
            const uint32_t t = t2 >> 1, rail = t2 & 1u;
            const u64 k = k0 + t;
            const bool live = t < TS && k < a.K;
            float acc = 0.0f;
            if (live) {
                const float* sv = rail ? s_vq : s_vi;
                const uint32_t j = t * sps + (N - 1) + (rail ? a.q_offset : 0u); /* local index of n_k (+ Q lag) */
                for (uint32_t i = 0; i < N; ++i) acc = mac<FMA>(acc, sv[rx_generic_skew(j - i)], s_taps[i]);
            }
            const float other = __shfl_xor_sync(0xffffffffu, acc, 1);
            if (live && rail == 0) {
                const float I = __fmul_rn(a.rx_gain, acc), Q = __fmul_rn(a.rx_gain, other);
                const uint32_t s = slice_point(s_slut + (k % a.n_tables) * a.n_const, a.n_const, I, Q);
                err += emit_symbol(a, f, k, s, I, Q);
                cmp += a.ref_bits ? a.bps : 0u;
            }
        }
    }
    block_count(a, err, cmp);
}

/*
 * Full-rate RX: materialises the (I,Q) stream the reference's Demodulator iterator yields
 * for every input sample (demodulator.rs:52-55).  2N MACs per sample: compute-bound, kept as
 * the parity / debug path.  One thread per output sample; tile = kThreads*4 samples.
 * dynamic smem: float s_vi[R]; float s_vq[R]; float s_taps[N]; R = tile + N - 1
 */
template <bool FMA>
__global__ void __launch_bounds__(kThreads) rx_fullrate_kernel(const __grid_constant__ RxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t TILE = kThreads * 4;
    const uint32_t N = a.n_taps, R = TILE + N - 1;
    float* s_vi = reinterpret_cast<float*>(smem_raw);
    float* s_vq = s_vi + R;
    float* s_taps = s_vq + R;
    for (uint32_t i = threadIdx.x; i < N; i += kThreads) s_taps[i] = a.taps[i];

    const u64 f = blockIdx.y;
    const float w = chan_w(a.ch, f), po = chan_po(a.ch, f);
    const u64 t0 = (u64)blockIdx.x * TILE;
    const long long nb = (long long)t0 - (long long)(N - 1);
    for (uint32_t j = threadIdx.x; j < R; j += kThreads) {
        long long n = nb + j;
        float2 v = make_float2(0.0f, 0.0f);
        if (n >= 0 && (u64)n < a.L) {
            float s, c;
            mg_sincosf_nco(__fadd_rn(nco_phase(w, a.sample0 + (u64)n), po), &s, &c);
            v = rx_mix(a, f, a.nz.frame0 + f, (u64)n, c, s);
        }
        s_vi[j] = v.x;
        s_vq[j] = v.y;
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < TILE; t += kThreads) {
        const u64 n = t0 + t;
        if (n >= a.L) break;
        const uint32_t j = t + N - 1;
        float ai = 0.0f, aq = 0.0f;
        for (uint32_t i = 0; i < N; ++i) {
            float c = s_taps[i];
            ai = mac<FMA>(ai, s_vi[j - i], c);
            aq = mac<FMA>(aq, s_vq[j - i], c);
        }
        a.filt[f * a.L + n] = make_float2(__fmul_rn(a.rx_gain, ai), __fmul_rn(a.rx_gain, aq));
    }
}

} /* namespace mg */
