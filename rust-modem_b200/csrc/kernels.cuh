/*
 * kernels.cuh -- sm_100a kernels of the batched modulate -> (AWGN) -> demodulate path.
 *
 * Reference lines each kernel replaces (paths relative to /root/reference/):
 *   tx_rect_kernel / tx_shaped_*   data.rs:66-79 (Bits), digital/<scheme>.rs i()/q() via a
 *                                  constellation LUT, carrier.rs:17-26 + util.rs:3-6 (NCO),
 *                                  modulator.rs:37-48,85-100 (mixer)
 *   rx_*_kernel                    demodulator.rs:44-55 + fir.rs:18-34 (two FIRs), plus the
 *                                  decimator / slicer / error-count extension
 *   awgn_kernel                    extension (Philox4x32-10 + Box-Muller)
 *
 * Arithmetic contract (what makes the output bit-identical to the scalar CPU path):
 *   - every binary32 operation the reference performs is issued as a separately rounded
 *     __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn (never contracted to FMA);
 *   - FIR sums run tap 0..N-1 in order inside one thread, starting from 0.0f;
 *   - sin/cos/log come from libm_f32.h (same algorithm + constants as glibc, in FP64).
 * Design: the NCO phase and its sin/cos depend only on the sample index, not on the frame,
 * so each CTA computes them ONCE for its tile of sample indices and then loops over many
 * frames; trig cost per sample is divided by the frames-per-block count.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "libm_f32.h"

namespace mg {

constexpr int kThreads = 256;
constexpr float kTwoPi = 6.28318548202514648437500f; /* (f32)PI * 2.0f, util.rs:4 */
constexpr int kMaxLut = 512;                          /* n_tables * 2^bps */
constexpr int kMaxFastTaps = 129;

typedef unsigned long long u64;

/* ------------------------------------------------------------------ small helpers */
struct ChannelView {
    const float* w;   /* per-channel sample_freq (nullable) */
    const float* po;  /* per-channel phase_offset (nullable) */
    float w0, po0;
    u64 frames_per_channel;
};
__device__ __forceinline__ float chan_w(const ChannelView& c, u64 f)
{
    return c.w ? __ldg(c.w + f / c.frames_per_channel) : c.w0;
}
__device__ __forceinline__ float chan_po(const ChannelView& c, u64 f)
{
    return c.po ? __ldg(c.po + f / c.frames_per_channel) : c.po0;
}

/* carrier.rs:17-19 + util.rs:3-6:  mod_trig(sample_freq * s as f32) */
__device__ __forceinline__ float nco_phase(float w, u64 s)
{
    float x = __fmul_rn(w, __ull2float_rn(s));
    float q = floorf(__fdiv_rn(x, kTwoPi));
    return __fsub_rn(x, __fmul_rn(kTwoPi, q));
}

/* modulator.rs:37-43 */
__device__ __forceinline__ float2 mix_iq(float i, float q, float c, float s)
{
    float2 r;
    r.x = __fsub_rn(__fmul_rn(i, c), __fmul_rn(q, s));
    r.y = __fadd_rn(__fmul_rn(i, s), __fmul_rn(q, c));
    return r;
}

template <bool FMA>
__device__ __forceinline__ float mac(float acc, float h, float c)
{
    /* fir.rs:23  s + history[cur] * coef */
    return FMA ? __fmaf_rn(h, c, acc) : __fadd_rn(acc, __fmul_rn(h, c));
}

/* digital/util.rs:5-11, MSB first */
__device__ __forceinline__ uint32_t pack_symbol(const uint8_t* p, uint32_t bps)
{
    uint32_t idx = 0;
    for (uint32_t j = 0; j < bps; ++j) idx = (idx << 1) | (__ldg(p + j) & 1u);
    return idx;
}

/* ------------------------------------------------------------------ Philox / AWGN */
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t r)
{
    return __double2float_rn(__fma_rn((double)r, 0x1p-32, 0x1p-33)); /* (2r+1)*2^-33: exact in FP64 */
}
/* n0 = rad*cos(theta); n1 (nullable) = rad*sin(theta) */
__device__ __forceinline__ void box_muller(uint32_t r0, uint32_t r1, float* n0, float* n1)
{
    float u1 = u01(r0), u2 = u01(r1);
    float rad = __fsqrt_rn(__fmul_rn(-2.0f, mg_logf_pos(u1)));
    float theta = __fmul_rn(kTwoPi, u2);
    float s, c;
    mg_sincosf(theta, &s, &c);
    *n0 = __fmul_rn(rad, c);
    if (n1) *n1 = __fmul_rn(rad, s);
}
struct Noise {
    float sigma; /* 0 => off */
    u64 seed, frame0;
};
/* real-part noise of sample n of global frame gf (the demodulator only reads .re) */
__device__ __forceinline__ float noise_re(const Noise& nz, u64 gf, u64 n)
{
    u64 pair = n >> 1;
    uint32_t r[4];
    philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)nz.seed,
                  (uint32_t)(nz.seed >> 32), r);
    float n0;
    if (n & 1) box_muller(r[2], r[3], &n0, nullptr);
    else box_muller(r[0], r[1], &n0, nullptr);
    return n0;
}

/* ================================================================== TX ============ */
struct TxArgs {
    const uint8_t* bits; /* [F][nbits] */
    u64 nbits;
    float2* tx;          /* [F][L] (nullable) */
    float2* iq;          /* [F][L] baseband (nullable) */
    u64 L, F, nsym;
    const float2* lut;   /* [n_tables][n_const] */
    uint32_t bps, sps, n_tables, n_const, q_offset;
    ChannelView ch;
    u64 sample0;
    uint32_t frames_per_block;
    /* shaped kernels */
    const float* taps;
    uint32_t n_taps;
    uint32_t sym_tile; /* symbols per CTA tile (generic shaped kernel) */
};

/* Symbol index of rail values at symbol m (EvenOddOffset semantics when q_offset != 0,
 * data.rs:102-122: cur[0] is replaced at the symbol edge, cur[1] half a symbol later). */
__device__ __forceinline__ uint32_t sym_index_plain(const uint8_t* fb, u64 m, uint32_t bps)
{
    return pack_symbol(fb + m * bps, bps);
}

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
            if (a.tx) __stcs(a.tx + o, mix_iq(ai, aq, cs.x, cs.y));
        }
    }
}

/*
 * Fast pulse-shaped TX for compile-time (SPS, NT): one thread owns one symbol period
 * (SPS consecutive samples).  The J = ceil(NT/SPS) symbols that reach it sit in registers,
 * the taps are kernel-parameter constants (constant-bank operands, no load instructions),
 * and the CTA's 2 KB-per-warp output is transposed through shared memory so that global
 * stores are full 128-bit coalesced.
 */
template <int NT>
struct TapsParam {
    float h[NT];
};

template <int SPS, int NT, bool FMA>
__global__ void __launch_bounds__(kThreads)
    tx_shaped_fast_kernel(const __grid_constant__ TxArgs a, const __grid_constant__ TapsParam<NT> taps)
{
    static_assert(SPS == 8, "output transpose below is written for 8 samples per symbol");
    constexpr int J = (NT + SPS - 1) / SPS; /* symbols reaching one output */
    constexpr int HALO = J - 1;
    __shared__ float2 s_lut[kMaxLut];
    __shared__ float2 s_sym[2][kThreads + HALO];
    __shared__ __align__(16) float4 s_out[kThreads / 32][32 * SPS / 2]; /* per warp: 32 symbols x 8 samples x 8 B */

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t i = tid; i < a.n_tables * a.n_const; i += kThreads) s_lut[i] = a.lut[i];

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);
    const u64 k0 = (u64)blockIdx.x * kThreads; /* first symbol of the tile */
    const u64 m = k0 + tid;                    /* this thread's symbol */

    float cs[SPS], sn[SPS];
#pragma unroll
    for (int p = 0; p < SPS; ++p) mg_sincosf(nco_phase(w, a.sample0 + m * SPS + p), &sn[p], &cs[p]);

    auto stage = [&](int buf, u64 f) {
        const uint8_t* fb = a.bits + f * a.nbits;
        for (int i = tid; i < kThreads + HALO; i += kThreads) {
            long long mm = (long long)k0 - HALO + i;
            float2 v = make_float2(0.0f, 0.0f);
            if (mm >= 0 && (u64)mm < a.nsym)
                v = s_lut[((u64)mm % a.n_tables) * a.n_const + sym_index_plain(fb, (u64)mm, a.bps)];
            s_sym[buf][i] = v;
        }
    };
    __syncthreads();
    if (f0 < f1) stage(0, f0);
    __syncthreads();

    int buf = 0;
    for (u64 f = f0; f < f1; ++f, buf ^= 1) {
        if (f + 1 < f1) stage(buf ^ 1, f + 1); /* overlap next frame's symbol fetch */
        float2 win[J];
#pragma unroll
        for (int j = 0; j < J; ++j) win[j] = s_sym[buf][tid + HALO - j]; /* symbol m - j */

        float4* wout = s_out[wid];
#pragma unroll
        for (int pp = 0; pp < SPS; pp += 2) {
            float2 o[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int p = pp + e;
                float ai = 0.0f, aq = 0.0f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    if (p + j * SPS < NT) {
                        ai = mac<FMA>(ai, win[j].x, taps.h[p + j * SPS]);
                        aq = mac<FMA>(aq, win[j].y, taps.h[p + j * SPS]);
                    }
                }
                o[e] = mix_iq(ai, aq, cs[p], sn[p]);
            }
            /* chunk c = pp/2 of row `lane` (4 chunks of 16 B per row), XOR-swizzled */
            const int c = pp >> 1;
            wout[lane * 4 + (c ^ ((lane >> 1) & 3))] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
        }
        __syncwarp();
        /* coalesced write-out of this warp's 32 symbols = 256 samples = 128 float4 */
        const u64 sym_w0 = k0 + (u64)wid * 32;
        float4* gout = reinterpret_cast<float4*>(a.tx + f * a.L + sym_w0 * SPS);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int g = lane + 32 * i; /* linear 16-byte chunk in the warp tile */
            const int row = g >> 2, c = g & 3;
            float4 v = wout[row * 4 + (c ^ ((row >> 1) & 3))];
            if (sym_w0 + row < a.nsym) __stcs(gout + g, v);
        }
        __syncthreads(); /* s_sym[buf^1] staged, s_out reusable */
    }
}

/* ================================================================== AWGN ========== */
/* In place: buf[f][n] += sigma * (n0 + j n1).  One thread = one Philox call = 2 samples. */
__global__ void __launch_bounds__(kThreads) awgn_kernel(float2* buf, u64 F, u64 L, Noise nz)
{
    const u64 pairs_per_frame = (L + 1) / 2;
    const u64 total = F * pairs_per_frame;
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 f = g / pairs_per_frame, pair = g % pairs_per_frame;
        const u64 gf = nz.frame0 + f;
        uint32_t r[4];
        philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)gf, (uint32_t)(gf >> 32),
                      (uint32_t)nz.seed, (uint32_t)(nz.seed >> 32), r);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const u64 n = pair * 2 + e;
            if (n >= L) break;
            float n0, n1;
            box_muller(r[2 * e], r[2 * e + 1], &n0, &n1);
            float2* p = buf + f * L + n;
            float2 v = *p;
            v.x = __fadd_rn(v.x, __fmul_rn(nz.sigma, n0));
            v.y = __fadd_rn(v.y, __fmul_rn(nz.sigma, n1));
            *p = v;
        }
    }
}

/* ================================================================== RX ============ */
struct RxArgs {
    const float2* rx; /* [F][L] */
    u64 L, F, K;      /* K decided symbols per frame */
    uint8_t* sym;     /* [F][K] nullable */
    uint8_t* bits;    /* [F][K*bps] nullable */
    float2* soft;     /* [F][K] nullable */
    float2* filt;     /* [F][L] nullable (full-rate kernel only) */
    const uint8_t* ref_bits; /* [F][ref_stride] nullable: count bit errors against these */
    u64 ref_stride;
    u64* counters;    /* [2]: errors, bits compared */
    const float2* slut; /* [n_tables][n_const] slicer table = slicer_gain * const_iq */
    uint32_t bps, sps, n_tables, n_const, q_offset, delay;
    float rx_gain;
    ChannelView ch;
    u64 sample0;
    uint32_t frames_per_block;
    const float* taps;
    uint32_t n_taps;
    uint32_t sym_tile;
    Noise nz;
};

/* extension 4: nearest point of the gain-scaled constellation, ties -> lowest index */
__device__ __forceinline__ uint32_t slice_point(const float2* t, uint32_t n, float I, float Q)
{
    uint32_t best = 0;
    float bd = 0.0f;
    for (uint32_t s = 0; s < n; ++s) {
        float2 c = t[s];
        float di = __fsub_rn(I, c.x), dq = __fsub_rn(Q, c.y);
        float d = __fadd_rn(__fmul_rn(di, di), __fmul_rn(dq, dq));
        if (s == 0 || d < bd) {
            bd = d;
            best = s;
        }
    }
    return best;
}

/* writes the decision of symbol k of frame f; returns the number of bit errors vs ref */
__device__ __forceinline__ uint32_t emit_symbol(const RxArgs& a, u64 f, u64 k, uint32_t s, float I, float Q)
{
    if (a.sym) a.sym[f * a.K + k] = (uint8_t)s;
    if (a.soft) a.soft[f * a.K + k] = make_float2(I, Q);
    if (a.bits) {
        uint8_t* o = a.bits + (f * a.K + k) * a.bps;
        for (uint32_t j = 0; j < a.bps; ++j) o[j] = (uint8_t)((s >> (a.bps - 1 - j)) & 1u);
    }
    uint32_t err = 0;
    if (a.ref_bits) {
        const uint8_t* r = a.ref_bits + f * a.ref_stride + k * a.bps;
        err = __popc(pack_symbol(r, a.bps) ^ s);
    }
    return err;
}

__device__ __forceinline__ void block_count(const RxArgs& a, uint32_t err, uint32_t nbits)
{
    __shared__ uint32_t s_err, s_cmp;
    if (threadIdx.x == 0) {
        s_err = 0;
        s_cmp = 0;
    }
    __syncthreads();
    err = __reduce_add_sync(0xffffffffu, err);
    nbits = __reduce_add_sync(0xffffffffu, nbits);
    if ((threadIdx.x & 31) == 0 && (err | nbits)) {
        atomicAdd(&s_err, err);
        atomicAdd(&s_cmp, nbits);
    }
    __syncthreads();
    if (threadIdx.x == 0 && a.counters && (s_err | s_cmp)) {
        atomicAdd(a.counters, (u64)s_err);
        atomicAdd(a.counters + 1, (u64)s_cmp);
    }
}

/* demodulator.rs:45-54: the two mixer products of sample n */
__device__ __forceinline__ float2 rx_mix(const RxArgs& a, const float2* frame, u64 gf, u64 n, float c, float s)
{
    float x = __ldcs(&frame[n].x);
    if (a.nz.sigma != 0.0f) x = __fadd_rn(x, __fmul_rn(a.nz.sigma, noise_re(a.nz, gf, n)));
    return make_float2(__fmul_rn(x, c), __fmul_rn(x, -s));
}

/*
 * Generic decimating RX (any sps / tap count / q_offset): one thread per symbol.
 * dynamic smem: float2 s_cs[R]; float s_vi[R]; float s_vq[R]; float s_taps[N]
 *   with R = (TS-1)*sps + N + q_offset samples per tile.
 */
template <bool FMA>
__global__ void __launch_bounds__(kThreads) rx_generic_kernel(const __grid_constant__ RxArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float2 s_slut[kMaxLut];
    const uint32_t TS = a.sym_tile, sps = a.sps, N = a.n_taps;
    const uint32_t R = (TS - 1) * sps + N + a.q_offset;
    float2* s_cs = reinterpret_cast<float2*>(smem_raw);
    float* s_vi = reinterpret_cast<float*>(s_cs + R);
    float* s_vq = s_vi + R;
    float* s_taps = s_vq + R;

    for (uint32_t i = threadIdx.x; i < a.n_tables * a.n_const; i += kThreads) s_slut[i] = a.slut[i];
    for (uint32_t i = threadIdx.x; i < N; i += kThreads) s_taps[i] = a.taps[i];

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0), po = chan_po(a.ch, f0);
    const u64 k0 = (u64)blockIdx.x * TS;
    /* sample index of s_v*[0]; may be negative (zero history, fir.rs:13) */
    const long long nb = (long long)(k0 * sps + a.delay) - (long long)(N - 1);
    for (uint32_t j = threadIdx.x; j < R; j += kThreads) {
        long long n = nb + j;
        float s = 0.0f, c = 0.0f;
        if (n >= 0 && (u64)n < a.L) mg_sincosf(__fadd_rn(nco_phase(w, a.sample0 + (u64)n), po), &s, &c);
        s_cs[j] = make_float2(c, s);
    }

    uint32_t err = 0, cmp = 0;
    for (u64 f = f0; f < f1; ++f) {
        const float2* frame = a.rx + f * a.L;
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < R; j += kThreads) {
            long long n = nb + j;
            float2 v = make_float2(0.0f, 0.0f);
            if (n >= 0 && (u64)n < a.L) {
                float2 cs = s_cs[j];
                v = rx_mix(a, frame, a.nz.frame0 + f, (u64)n, cs.x, cs.y);
            }
            s_vi[j] = v.x;
            s_vq[j] = v.y;
        }
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < TS; t += kThreads) {
            const u64 k = k0 + t;
            if (k >= a.K) break;
            const uint32_t ji = t * sps + (N - 1); /* local index of n_k */
            const uint32_t jq = ji + a.q_offset;
            float ai = 0.0f, aq = 0.0f;
            for (uint32_t i = 0; i < N; ++i) {
                float c = s_taps[i];
                ai = mac<FMA>(ai, s_vi[ji - i], c);
                aq = mac<FMA>(aq, s_vq[jq - i], c);
            }
            const float I = __fmul_rn(a.rx_gain, ai), Q = __fmul_rn(a.rx_gain, aq);
            const uint32_t s = slice_point(s_slut + (k % a.n_tables) * a.n_const, a.n_const, I, Q);
            err += emit_symbol(a, f, k, s, I, Q);
            cmp += a.ref_bits ? a.bps : 0u;
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
    const float2* frame = a.rx + f * a.L;
    const u64 t0 = (u64)blockIdx.x * TILE;
    const long long nb = (long long)t0 - (long long)(N - 1);
    for (uint32_t j = threadIdx.x; j < R; j += kThreads) {
        long long n = nb + j;
        float2 v = make_float2(0.0f, 0.0f);
        if (n >= 0 && (u64)n < a.L) {
            float s, c;
            mg_sincosf(__fadd_rn(nco_phase(w, a.sample0 + (u64)n), po), &s, &c);
            v = rx_mix(a, frame, a.nz.frame0 + f, (u64)n, c, s);
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

/*
 * Fast decimating RX for compile-time (SPS = 8, NT taps): the hot kernel of the loopback.
 *
 *  phase A  every thread loads 128-bit pairs of complex samples of the tile + (NT-1) halo,
 *           multiplies .re by the CTA-resident (cos, -sin) table and stores the two mixer
 *           products into padded shared memory;
 *  phase B  every thread owns R = 2 consecutive symbols and walks taps 0..NT-1 in order;
 *           taps are constant-bank operands, samples slide through a 2x8-register window
 *           fed by conflict-free 128-bit shared loads (one new 8-sample block serves 16 MACs
 *           per rail);
 *  phase C  slice, pack, count errors, coalesced stores.
 *
 * smem layout: sample j of the tile lives at pos(j) = j' + 4*(j' >> 5), j' = j + SHIFT,
 * i.e. 4 floats of padding per 32, which makes the stride-16-float 128-bit loads of phase B
 * hit 8 distinct bank groups per quarter warp.
 */
template <int SPS, int NT>
struct RxFastCfg {
    static constexpr int R = 2;                          /* symbols per thread */
    static constexpr int TS = kThreads * R;              /* symbols per tile */
    static constexpr int NSAMP = (TS - 1) * SPS + NT;    /* samples staged per tile */
    static constexpr int SHIFT = (4 - ((NT - SPS) & 3)) & 3; /* makes every 8-sample block 16B aligned */
    static constexpr int NPOS = NSAMP + SHIFT;
    static constexpr int PADDED = NPOS + 4 * ((NPOS + 31) / 32) + 4;
    static constexpr int NPAIR = (NSAMP + 2) / 2 + 1;    /* 128-bit global loads per tile (may start one sample early) */
    static constexpr size_t SMEM = sizeof(float) * 2 * PADDED + sizeof(float4) * NPAIR;
};

__device__ __forceinline__ int rx_pos(int jp) { return jp + 4 * (jp >> 5); }

template <int SPS, int NT, bool FMA>
__global__ void __launch_bounds__(kThreads, 2)
    rx_fast_kernel(const __grid_constant__ RxArgs a, const __grid_constant__ TapsParam<NT> taps)
{
    static_assert(SPS == 8, "window logic below assumes 8 samples per symbol");
    using C = RxFastCfg<SPS, NT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float2 s_slut[kMaxLut];
    float* s_vi = reinterpret_cast<float*>(smem_raw);
    float* s_vq = s_vi + C::PADDED;
    float4* s_cs = reinterpret_cast<float4*>(s_vq + C::PADDED); /* per pair: (c0, -s0, c1, -s1) */

    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < a.n_tables * a.n_const; i += kThreads) s_slut[i] = a.slut[i];

    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0), po = chan_po(a.ch, f0);
    const u64 k0 = (u64)blockIdx.x * C::TS;
    /* sample index of tile-local j = 0 (may be negative) */
    const long long nb = (long long)(k0 * SPS + a.delay) - (long long)(NT - 1);
    /* global loads are 16-byte pairs (n even); pair q covers samples nb_even + 2q, +1 */
    const long long nb_even = nb & ~1LL; /* floor to even (two's complement: also right for negatives) */
    for (int q = tid; q < C::NPAIR; q += kThreads) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            long long n = nb_even + 2 * q + e;
            float s = 0.0f, c = 0.0f;
            if (n >= 0 && (u64)n < a.L) mg_sincosf(__fadd_rn(nco_phase(w, a.sample0 + (u64)n), po), &s, &c);
            v[2 * e] = c;
            v[2 * e + 1] = -s;
        }
        s_cs[q] = make_float4(v[0], v[1], v[2], v[3]);
    }

    const u64 ka = k0 + 2 * (u64)tid; /* this thread's two symbols: ka, ka+1 */
    const uint32_t toff0 = (uint32_t)(ka % a.n_tables) * a.n_const;
    const uint32_t toff1 = (uint32_t)((ka + 1) % a.n_tables) * a.n_const;
    uint32_t err = 0, cmp = 0;
    const bool frame_pairs_ok = (a.L & 1) == 0; /* every frame starts 16B aligned */
    for (u64 f = f0; f < f1; ++f) {
        const float2* frame = a.rx + f * a.L;
        const u64 gf = a.nz.frame0 + f;
        __syncthreads(); /* previous frame's phase B finished; s_cs / s_slut visible */
        /* ---- phase A: issue every global load of the tile first (memory-level parallelism) */
        constexpr int ITER = (C::NPAIR + kThreads - 1) / kThreads;
        float x0[ITER], x1[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int q = tid + it * kThreads;
            const long long n = nb_even + 2 * q;
            x0[it] = 0.0f;
            x1[it] = 0.0f;
            if (q < C::NPAIR) {
                if (n >= 0 && (u64)n + 1 < a.L && frame_pairs_ok) {
                    float4 t = __ldcs(reinterpret_cast<const float4*>(frame + n));
                    x0[it] = t.x;
                    x1[it] = t.z;
                } else {
                    if (n >= 0 && (u64)n < a.L) x0[it] = __ldcs(&frame[n].x);
                    if (n + 1 >= 0 && (u64)(n + 1) < a.L) x1[it] = __ldcs(&frame[n + 1].x);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int q = tid + it * kThreads;
            if (q >= C::NPAIR) break;
            const long long n = nb_even + 2 * q;
            float v0 = x0[it], v1 = x1[it];
            if (a.nz.sigma != 0.0f) {
                if (n >= 0 && (u64)n < a.L) v0 = __fadd_rn(v0, __fmul_rn(a.nz.sigma, noise_re(a.nz, gf, (u64)n)));
                if (n + 1 >= 0 && (u64)(n + 1) < a.L)
                    v1 = __fadd_rn(v1, __fmul_rn(a.nz.sigma, noise_re(a.nz, gf, (u64)(n + 1))));
            }
            const float4 cs = s_cs[q];
            const int j0 = (int)(n - nb); /* tile-local index of the pair's first sample: -1 or >= 0 */
            if (j0 >= 0 && j0 < C::NSAMP) {
                const int p = rx_pos(j0 + C::SHIFT);
                s_vi[p] = __fmul_rn(v0, cs.x);
                s_vq[p] = __fmul_rn(v0, cs.y);
            }
            if (j0 + 1 >= 0 && j0 + 1 < C::NSAMP) {
                const int p = rx_pos(j0 + 1 + C::SHIFT);
                s_vi[p] = __fmul_rn(v1, cs.z);
                s_vq[p] = __fmul_rn(v1, cs.w);
            }
        }
        __syncthreads();
        /* ---- phase B: symbols r0 = 2*tid, r0+1; local index of n_k for symbol r is NT-1 + 8r */
        {
            const int base = NT - 1 + 16 * tid + C::SHIFT; /* shifted index of symbol r0's instant */
            float ai0 = 0.0f, aq0 = 0.0f, ai1 = 0.0f, aq1 = 0.0f;
            float pi[8], pq[8], ci[8], cq[8];
            /* prev = block "-1" = shifted indices [base+1, base+8] */
            {
                const int g = base + 1;
                float4 t0 = *reinterpret_cast<const float4*>(s_vi + rx_pos(g));
                float4 t1 = *reinterpret_cast<const float4*>(s_vi + rx_pos(g + 4));
                float4 u0 = *reinterpret_cast<const float4*>(s_vq + rx_pos(g));
                float4 u1 = *reinterpret_cast<const float4*>(s_vq + rx_pos(g + 4));
                pi[0] = t0.x; pi[1] = t0.y; pi[2] = t0.z; pi[3] = t0.w; pi[4] = t1.x; pi[5] = t1.y; pi[6] = t1.z; pi[7] = t1.w;
                pq[0] = u0.x; pq[1] = u0.y; pq[2] = u0.z; pq[3] = u0.w; pq[4] = u1.x; pq[5] = u1.y; pq[6] = u1.z; pq[7] = u1.w;
            }
            constexpr int NB = (NT + 7) / 8;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                /* block b = shifted indices [base-8b-7, base-8b]; element e <-> index base-8b-7+e */
                const int g = base - 8 * b - 7;
                {
                    /* for the last partial block g may dip below 0 only when NT % 8 != 0 and the
                       block holds just tap 8b (index base-8b): guard the low half */
                    if (g >= 0) {
                        float4 t0 = *reinterpret_cast<const float4*>(s_vi + rx_pos(g));
                        float4 u0 = *reinterpret_cast<const float4*>(s_vq + rx_pos(g));
                        ci[0] = t0.x; ci[1] = t0.y; ci[2] = t0.z; ci[3] = t0.w;
                        cq[0] = u0.x; cq[1] = u0.y; cq[2] = u0.z; cq[3] = u0.w;
                    } else {
                        ci[0] = ci[1] = ci[2] = ci[3] = 0.0f;
                        cq[0] = cq[1] = cq[2] = cq[3] = 0.0f;
                    }
                    float4 t1 = *reinterpret_cast<const float4*>(s_vi + rx_pos(g + 4));
                    float4 u1 = *reinterpret_cast<const float4*>(s_vq + rx_pos(g + 4));
                    ci[4] = t1.x; ci[5] = t1.y; ci[6] = t1.z; ci[7] = t1.w;
                    cq[4] = u1.x; cq[5] = u1.y; cq[6] = u1.z; cq[7] = u1.w;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = 8 * b + u; /* tap index */
                    if (i < NT) {
                        const float c = taps.h[i];
                        ai0 = mac<FMA>(ai0, ci[7 - u], c);
                        aq0 = mac<FMA>(aq0, cq[7 - u], c);
                        ai1 = mac<FMA>(ai1, pi[7 - u], c);
                        aq1 = mac<FMA>(aq1, pq[7 - u], c);
                    }
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    pi[e] = ci[e];
                    pq[e] = cq[e];
                }
            }
            /* ---- phase C */
            const float I0 = __fmul_rn(a.rx_gain, ai0), Q0 = __fmul_rn(a.rx_gain, aq0);
            const float I1 = __fmul_rn(a.rx_gain, ai1), Q1 = __fmul_rn(a.rx_gain, aq1);
            if (ka < a.K) {
                const uint32_t s0 = slice_point(s_slut + toff0, a.n_const, I0, Q0);
                err += emit_symbol(a, f, ka, s0, I0, Q0);
                cmp += a.ref_bits ? a.bps : 0u;
            }
            if (ka + 1 < a.K) {
                const uint32_t s1 = slice_point(s_slut + toff1, a.n_const, I1, Q1);
                err += emit_symbol(a, f, ka + 1, s1, I1, Q1);
                cmp += a.ref_bits ? a.bps : 0u;
            }
        }
    }
    block_count(a, err, cmp);
}

} /* namespace mg */
