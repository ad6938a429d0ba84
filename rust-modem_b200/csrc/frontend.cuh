/*
 * frontend.cuh -- the kernels either side of the memoryless hot path:
 *
 *   tone_kernel          modulate.rs:118-126: Modulator + phasor::Raw (modulator.rs:51-62, phasor.rs:5-24)
 *   phasor_scan_kernel   DigitalPhasor::update of the stateful schemes (bfsk.rs:43-55, mfsk.rs:68-75,
 *                        dmpsk.rs:29-33): the symbol-to-symbol phase recurrence, one lane per frame
 *   tx_phasor_kernel     i()/q() of bfsk / mfsk / cpfsk / msk / dmpsk + Carrier + IQSample::modulate
 *   lock_phase_kernel    Demodulator::lock_phase (demodulator.rs:32-36) over the analytic signal of
 *                        demodulate.rs:31-34 (Hilbert FIR) with PLL::handle (pll.rs:16-22)
 *
 * Arithmetic contract as in common.cuh: every binary32 operation of the reference is one separately
 * rounded __f*_rn, in the reference's order; sin/cos/atan2 come from libm_f32.h (bit-identical to glibc).
 */
#pragma once

#include "common.cuh"

namespace mg {

constexpr float kPi = 3.14159274101257324218750f; /* std::f32::consts::PI */

/* util.rs:3-6 */
__device__ __forceinline__ float mod_trig(float x)
{
    const float q = floorf(__fdiv_rn(x, kTwoPi));
    return __fsub_rn(x, __fmul_rn(kTwoPi, q));
}

/* ------------------------------------------------------------------ preamble tone */
/* sample j of frame f: IQSample{carrier: phase(sample0 + j), i: A, q: 0.0}.modulate() (modulator.rs:37-48).
 * tx (nullable) [F][n] complex; re (nullable) re[f * re_stride + j]. */
__global__ void __launch_bounds__(kThreads)
    tone_kernel(float2* tx, float* re, u64 re_stride, u64 F, u64 n, float amplitude, ChannelView ch, u64 sample0)
{
    const u64 total = F * n;
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 f = g / n, j = g % n;
        float sn, cs;
        mg_sincosf(nco_phase(chan_w(ch, f), sample0 + j), &sn, &cs);
        const float2 v = mix_iq(amplitude, 0.0f, cs, sn); /* phasor.rs:17-18: i = amplitude, q = 0.0 */
        if (tx) tx[g] = v;
        if (re) re[f * re_stride + j] = v.x;
    }
}

/* ------------------------------------------------------------------ stateful phasors */
struct PhasorArgs {
    uint32_t kind;          /* MODEM_PHASOR_* of include/modem_gpu.h */
    uint32_t bps;
    uint32_t increase_map;  /* mfsk.rs:31-35 vs :22-27 */
    int max_symbol;         /* mfsk.rs:18 */
    float amplitude, deviation, phase, shift;
    float samples_per_bit;  /* msk.rs:17 as f32 (msk.rs:22) */
    const float* state;     /* [F][nsym]: the phasor's phase while symbol k is held (bfsk / mfsk / dmpsk) */
};
enum { kPhTable = 0, kPhBfsk = 1, kPhMfsk = 2, kPhCpfsk = 3, kPhMsk = 4, kPhDmpsk = 5 };

__device__ __forceinline__ float mfsk_coef(const PhasorArgs& p, uint32_t sym)
{
    if (p.increase_map) return (float)(uint8_t)(2u * sym); /* mfsk.rs:33: (2 * symbol) as f32, u8 arithmetic */
    return (float)(2 * (int)sym - p.max_symbol);           /* mfsk.rs:25 */
}

/*
 * The recurrence DigitalModulator::next drives through DigitalPhasor::update at every symbol edge
 * (modulator.rs:90-93: Changed => update(carrier.sample, bits), carrier.sample already incremented, so symbol
 * k of a frame updates with s = sample0 + k*sps + 1).  binary32 addition is not associative, so the chain is
 * evaluated in order, one lane per frame; state[f][k] = the phase in force while symbol k is held.  Four
 * values are buffered per 128-bit store.
 */
__global__ void __launch_bounds__(128)
    phasor_scan_kernel(const uint8_t* bits, u64 nbits, u64 F, u64 nsym, uint32_t sps, u64 sample0, PhasorArgs p, float* state)
{
    const u64 f = (u64)blockIdx.x * 128 + threadIdx.x;
    if (f >= F) return;
    const uint8_t* fb = bits + f * nbits;
    float* out = state + f * nsym;
    const bool vec = ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    float phase = p.kind == kPhDmpsk ? p.phase : 0.0f; /* dmpsk.rs:20; bfsk.rs:18; mfsk.rs:55 */
    float cur_coef = 0.0f;                             /* mfsk.rs:56 */
    uint32_t prev = 0;                                 /* bfsk.rs:19 */
    float buf[4];
    for (u64 k = 0; k < nsym; ++k) {
        const u64 s = sample0 + k * sps + 1;
        const uint32_t sym = pack_symbol(fb + k * p.bps, p.bps);
        if (p.kind == kPhBfsk) { /* bfsk.rs:43-55 */
            if (sym != prev) {
                /* rads(s, 1) = 1 as f32 * deviation * s as f32 (bfsk.rs:27-29) */
                const float d = sym == 1 ? -__fmul_rn(__fmul_rn(1.0f, p.deviation), __ull2float_rn(s))
                                         : __fmul_rn(__fmul_rn(1.0f, p.deviation), __ull2float_rn(s - 1));
                phase = mod_trig(__fadd_rn(phase, d));
                prev = sym;
            }
        } else if (p.kind == kPhMfsk) { /* mfsk.rs:68-75 */
            const float next = mfsk_coef(p, sym);
            phase = __fadd_rn(phase, __fmul_rn(__fmul_rn(__fsub_rn(cur_coef, next), p.deviation), __ull2float_rn(s)));
            phase = mod_trig(phase);
            cur_coef = next;
        } else { /* dmpsk.rs:29-33 */
            phase = mod_trig(__fadd_rn(phase, __fmul_rn((float)sym, p.shift)));
        }
        buf[k & 3] = phase;
        if (vec && (k & 3) == 3) {
            *reinterpret_cast<float4*>(out + k - 3) = make_float4(buf[0], buf[1], buf[2], buf[3]);
        } else if (!vec) {
            out[k] = phase;
        }
    }
    if (vec)
        for (u64 k = nsym & ~3ull; k < nsym; ++k) out[k] = buf[k & 3];
}

/*
 * Per-sample part.  One thread owns U sample indices of the tile (stride kThreads, coalesced stores) and
 * loops over the CTA's frames: the carrier's (cos, sin) and, for msk, the phasor's own (cos, sin) depend on
 * the sample index only and are evaluated once per thread; bfsk / mfsk / cpfsk need one sincos per sample and
 * frame (their argument depends on the frame's bits), dmpsk one per symbol.
 */
template <int KIND>
__global__ void __launch_bounds__(kThreads) tx_phasor_kernel(const __grid_constant__ TxArgs a, const __grid_constant__ PhasorArgs p)
{
    constexpr int U = 4;
    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);

    u64 n[U];
    float cs[U], sn[U], sf[U], pc[U], ps[U];
    uint32_t ki[U], kq[U];
    bool qv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        n[u] = ((u64)blockIdx.x * U + u) * kThreads + threadIdx.x;
        ki[u] = (uint32_t)(n[u] / a.sps);
        qv[u] = n[u] >= a.q_offset;
        kq[u] = qv[u] ? (uint32_t)((n[u] - a.q_offset) / a.sps) : 0u;
        mg_sincosf(nco_phase(w, a.sample0 + n[u]), &sn[u], &cs[u]);
        sf[u] = __ull2float_rn(a.sample0 + n[u] + 1); /* the phasor's `s as f32` (modulator.rs:86-97) */
        pc[u] = ps[u] = 0.0f;
        if (KIND == kPhMsk) /* msk.rs:21-23: PI / 2.0 * s as f32 / samples_per_bit as f32 */
            mg_sincosf(__fdiv_rn(__fmul_rn(kPi / 2.0f, sf[u]), p.samples_per_bit), &ps[u], &pc[u]);
    }

    for (u64 f = f0; f < f1; ++f) {
        const uint8_t* fb = a.bits + f * a.nbits;
        const float* st = p.state + f * a.nsym;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (n[u] >= a.L) continue;
            float bi, bq;
            if (KIND == kPhMsk) { /* msk.rs:29-35, bits through EvenOddOffset (data.rs:102-122) */
                const uint32_t b0 = __ldg(fb + (u64)ki[u] * 2) & 1u;
                const uint32_t b1 = qv[u] ? (__ldg(fb + (u64)kq[u] * 2 + 1) & 1u) : 0u;
                bi = __fmul_rn(__fmul_rn(p.amplitude, (float)(2 * (int)b0 - 1)), pc[u]);
                bq = __fmul_rn(__fmul_rn(-p.amplitude, (float)(2 * (int)b1 - 1)), ps[u]);
            } else {
                const uint32_t sym = pack_symbol(fb + (u64)ki[u] * p.bps, p.bps);
                float inner;
                if (KIND == kPhBfsk) /* bfsk.rs:23-29: b as f32 * deviation * s as f32 + phase */
                    inner = __fadd_rn(__fmul_rn(__fmul_rn((float)sym, p.deviation), sf[u]), __ldg(st + ki[u]));
                else if (KIND == kPhMfsk) /* mfsk.rs:60-62 */
                    inner = __fadd_rn(__fmul_rn(__fmul_rn(mfsk_coef(p, sym), p.deviation), sf[u]), __ldg(st + ki[u]));
                else if (KIND == kPhCpfsk) /* cpfsk.rs:25-31: coef = 2.0 * symbol as f32 */
                    inner = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, (float)sym), p.deviation), sf[u]);
                else /* dmpsk.rs:35-41 */
                    inner = __ldg(st + ki[u]);
                float s_, c_;
                mg_sincosf(inner, &s_, &c_);
                bi = __fmul_rn(p.amplitude, c_);
                bq = __fmul_rn(p.amplitude, s_);
            }
            const u64 o = f * a.L + n[u];
            if (a.iq) __stcs(a.iq + o, make_float2(bi, bq));
            if (a.tx || a.re) {
                const float2 m = mix_iq(bi, bq, cs[u], sn[u]);
                if (a.tx) __stcs(a.tx + o, m);
                if (a.re) __stcs(a.re + f * a.re_stride + a.re_offset + n[u], m.x);
            }
        }
    }
}

/* ------------------------------------------------------------------ PLL phase lock */
struct LockArgs {
    const void* samples; /* [F][stride] in fmt */
    uint32_t fmt;        /* 0 complex f32 (analytic supplied), 1 real f32, 2 real i16 */
    u64 F, stride;
    const float* htaps;  /* Hilbert FIR (demodulate.rs:48-72) for the real formats */
    uint32_t n_h;
    uint32_t lock;       /* LOCK_SAMPLES (demodulator.rs:5) */
    ChannelView ch;
    u64 sample0;
    float* po;           /* [F] out: PLL.phase_offset after the lock */
};

__device__ __forceinline__ float lock_sample(const LockArgs& a, u64 f, u64 t)
{
    const u64 o = f * a.stride + t;
    if (a.fmt == 0) return __ldg(&reinterpret_cast<const float2*>(a.samples)[o].x);
    if (a.fmt == 1) return __ldg(reinterpret_cast<const float*>(a.samples) + o);
    return (float)__ldg(reinterpret_cast<const short*>(a.samples) + o);
}

/* 64 strictly sequential PLL steps per frame (each feeds the next through phase_offset): one lane per frame. */
__global__ void __launch_bounds__(64) lock_phase_kernel(const __grid_constant__ LockArgs a)
{
    const u64 f = (u64)blockIdx.x * 64 + threadIdx.x;
    if (f >= a.F) return;
    const float CHANGE = 0.447214f; /* pll.rs:3 */
    const float w = chan_w(a.ch, f);
    float po = 0.0f; /* pll.rs:10-14 */
    for (uint32_t t = 0; t < a.lock; ++t) {
        const float x_re = lock_sample(a, f, t);
        float x_im;
        if (a.fmt == 0) {
            x_im = __ldg(&reinterpret_cast<const float2*>(a.samples)[f * a.stride + t].y);
        } else {
            /* hfir.add(x) (demodulate.rs:33, fir.rs:18-34): fold over coefs from 0.0, newest sample first,
             * zero history before the frame starts */
            float s = 0.0f;
            for (uint32_t k = 0; k < a.n_h; ++k) {
                const float h = t >= k ? lock_sample(a, f, t - k) : 0.0f;
                s = __fadd_rn(s, __fmul_rn(h, __ldg(a.htaps + k)));
            }
            x_im = s;
        }
        /* pll.rs:16-22 */
        const float inner = __fadd_rn(nco_phase(w, a.sample0 + t), po);
        float sn, cs;
        mg_sincosf(inner, &sn, &cs);
        const float c_re = cs, c_im = -sn; /* carrier.conj() */
        const float m_re = __fsub_rn(__fmul_rn(x_re, c_re), __fmul_rn(x_im, c_im)); /* num-0.1.35 Complex Mul */
        const float m_im = __fadd_rn(__fmul_rn(x_re, c_im), __fmul_rn(x_im, c_re));
        const float err = mg_atan2f(m_im, m_re);                                    /* arg() = im.atan2(re) */
        po = __fadd_rn(po, __fmul_rn(CHANGE, err));
    }
    a.po[f] = po;
}

} /* namespace mg */
