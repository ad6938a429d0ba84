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
    const float2* siq;      /* [F][nsym]: dmpsk's (i, q) per symbol (phasor_symbol_iq_kernel) */
    const float2* cp_tab;   /* [2^bps][L]: cpfsk's (i, q) per (symbol value, sample) (nullable: evaluate per sample) */
};
enum { kPhTable = 0, kPhBfsk = 1, kPhMfsk = 2, kPhCpfsk = 3, kPhMsk = 4, kPhDmpsk = 5 };

__device__ __forceinline__ float mfsk_coef(const PhasorArgs& p, uint32_t sym)
{
    if (p.increase_map) return (float)(uint8_t)(2u * sym); /* mfsk.rs:33: (2 * symbol) as f32, u8 arithmetic */
    return (float)(2 * (int)sym - p.max_symbol);           /* mfsk.rs:25 */
}

/* x / TWO_PI, IEEE round-to-nearest, without the division unit on the dependent path: q0 = x*r, one exact
 * residual e = x - q0*d, q = q0 + e*r with r = fl(1/d).  tools/check_div.c compares this against x / d for EVERY
 * binary32 x: identical for 1e-30 <= |x| <= 1e30 (0 mismatches of 1.7e9); outside that range (signed zeros,
 * subnormal quotients, overflow) the IEEE division is used. */
__device__ __forceinline__ float div_two_pi(float x)
{
    const float ax = fabsf(x);
    if (ax >= 1e-30f && ax <= 1e30f) {
        const float r = 0.15915493667125701904296875f; /* fl(1 / kTwoPi) */
        const float q0 = __fmul_rn(x, r);
        const float e = __fmaf_rn(-q0, kTwoPi, x);
        return __fmaf_rn(e, r, q0);
    }
    return __fdiv_rn(x, kTwoPi);
}
__device__ __forceinline__ float mod_trig_fast(float x) /* == mod_trig(x) bit for bit */
{
    return __fsub_rn(x, __fmul_rn(kTwoPi, floorf(div_two_pi(x))));
}

/*
 * The recurrence DigitalModulator::next drives through DigitalPhasor::update at every symbol edge
 * (modulator.rs:90-93: Changed => update(carrier.sample, bits), carrier.sample already incremented, so symbol
 * k of a frame updates with s = sample0 + k*sps + 1).  binary32 addition is not associative, so the chain is
 * evaluated in order, one lane per frame; state[f][k] = the phase in force while symbol k is held.
 *
 * One CTA = 32 frames = two warps.  The frames' bit rows and state rows are far apart in memory, so a lane
 * walking its own row in global memory would touch 32 different lines per access and wait ~800 cycles for each.
 * Instead a chunk of SCH symbols of all 32 frames goes through shared memory: the HELPER warp copies the next
 * chunk's bit bytes in with cp.async (coalesced row segments) and writes the previous chunk's phases out
 * (coalesced row segments) while the WALKER warp -- lane = frame -- runs the recurrence over the current chunk
 * out of its own shared row (row stride an odd number of words: conflict-free).  The walker's loop carries
 * only the dependent chain; the measured single-warp version spent 6x longer on staging than on the chain.
 */
constexpr int kScanChunk = 64; /* symbols per chunk */

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

/* dynamic shared memory: uint8_t s_bits[2][32][SCH*bps + 4], then float s_state[2][32][SCH + 1] */
template <int KIND>
__global__ void __launch_bounds__(64, 4)
    phasor_scan_kernel(const uint8_t* bits, u64 nbits, u64 F, u64 nsym, uint32_t sps, u64 sample0, PhasorArgs p, float* state,
                       int aligned4)
{
    constexpr int SCH = kScanChunk;
    extern __shared__ __align__(16) unsigned char scan_smem[];
    const uint32_t bps = p.bps;
    const uint32_t rowb = SCH * bps + 4; /* bytes per staged bit row: 16*bps + 1 words */
    uint8_t* s_bits = scan_smem;
    float(*s_state)[32][SCH + 1] = reinterpret_cast<float(*)[32][SCH + 1]>(scan_smem + 2 * 32 * rowb);
    const int lane = threadIdx.x & 31;
    const bool walker = threadIdx.x < 32;
    const u64 f0 = (u64)blockIdx.x * 32;
    const uint32_t rows = (uint32_t)min((u64)32, F - f0);
    const u64 f = f0 + lane;
    const u64 nchunks = (nsym + SCH - 1) / SCH;

    float phase = KIND == kPhDmpsk ? p.phase : 0.0f; /* dmpsk.rs:20; bfsk.rs:18; mfsk.rs:55 */
    float cur_coef = 0.0f;                           /* mfsk.rs:56 */
    uint32_t prev = 0;                               /* bfsk.rs:19 */
    u64 s = sample0 + 1;                             /* Carrier.sample seen by update() at symbol 0 */

    /* helper: bit bytes of symbols [k0, k0 + SCH) of every row -> s_bits[buf] */
    auto stage = [&](int buf, u64 k0) {
        uint8_t* dst = s_bits + (size_t)buf * 32 * rowb;
        const uint32_t nb = (uint32_t)min((u64)SCH, nsym - k0) * bps; /* valid bytes of the segment */
        const uint8_t* src = bits + f0 * nbits + k0 * bps;
        for (uint32_t r = 0; r < rows; ++r, src += nbits, dst += rowb) {
            if (aligned4) {
                for (uint32_t w = 4 * lane; w < nb; w += 128) {
                    if (w + 4 <= nb) cp_async4(dst + w, src + w);
                    else
                        for (uint32_t e = w; e < nb; ++e) dst[e] = __ldg(src + e);
                }
            } else {
                for (uint32_t e = lane; e < nb; e += 32) dst[e] = __ldg(src + e);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    };
    /* helper: phases of chunk c (s_state[buf]) -> global, coalesced row segments */
    auto flush = [&](int buf, u64 k0) {
        const uint32_t cnt = (uint32_t)min((u64)SCH, nsym - k0);
        float* dst = state + f0 * nsym + k0;
        for (uint32_t r = 0; r < rows; ++r, dst += nsym)
#pragma unroll
            for (int h = 0; h < SCH / 32; ++h) {
                const uint32_t j = lane + 32 * h;
                if (j < cnt) dst[j] = s_state[buf][r][j];
            }
    };

    if (!walker) stage(0, 0);
    __syncthreads();
    for (u64 c = 0; c < nchunks; ++c) {
        const int buf = (int)(c & 1);
        const u64 k0 = c * SCH;
        if (walker) {
            const uint32_t cnt = (uint32_t)min((u64)SCH, nsym - k0);
            const uint8_t* row = s_bits + (size_t)buf * 32 * rowb + lane * rowb;
            if (f < F) {
#pragma unroll 4
                for (uint32_t j = 0; j < cnt; ++j, s += sps) {
                    uint32_t sym = 0;
                    for (uint32_t e = 0; e < bps; ++e) sym = (sym << 1) | (row[j * bps + e] & 1u); /* digital/util.rs:5-11 */
                    if (KIND == kPhBfsk) { /* bfsk.rs:43-55 */
                        if (sym != prev) {
                            /* rads(s, 1) = 1 as f32 * deviation * s as f32 (bfsk.rs:27-29) */
                            const float d = sym == 1 ? -__fmul_rn(__fmul_rn(1.0f, p.deviation), __ull2float_rn(s))
                                                     : __fmul_rn(__fmul_rn(1.0f, p.deviation), __ull2float_rn(s - 1));
                            phase = mod_trig_fast(__fadd_rn(phase, d));
                            prev = sym;
                        }
                    } else if (KIND == kPhMfsk) { /* mfsk.rs:68-75 */
                        const float next = mfsk_coef(p, sym);
                        phase = __fadd_rn(phase, __fmul_rn(__fmul_rn(__fsub_rn(cur_coef, next), p.deviation), __ull2float_rn(s)));
                        phase = mod_trig_fast(phase);
                        cur_coef = next;
                    } else { /* dmpsk.rs:29-33 */
                        phase = mod_trig_fast(__fadd_rn(phase, __fmul_rn((float)sym, p.shift)));
                    }
                    s_state[buf][lane][j] = phase;
                }
            }
        } else {
            if (c > 0) flush(buf ^ 1, k0 - SCH);
            if (c + 1 < nchunks) stage(buf ^ 1, k0 + SCH);
        }
        __syncthreads();
    }
    if (!walker && nchunks) flush((int)((nchunks - 1) & 1), (nchunks - 1) * SCH);
}

/* dmpsk.rs:35-41: (amplitude * cos(phase), amplitude * sin(phase)) once per symbol instead of once per sample */
__global__ void __launch_bounds__(kThreads) phasor_symbol_iq_kernel(const float* state, u64 n, float amplitude, float2* siq)
{
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < n; g += (u64)gridDim.x * kThreads) {
        float sn, cs;
        mg_sincosf(__ldg(state + g), &sn, &cs);
        siq[g] = make_float2(__fmul_rn(amplitude, cs), __fmul_rn(amplitude, sn));
    }
}

/* cpfsk.rs:25-43: i/q depend on (symbol value, sample counter) only -- the same for every frame.  Table row v,
 * entry n = (A cos, A sin) of (2.0 * v as f32) * freq * (sample0 + n + 1) as f32. */
__global__ void __launch_bounds__(kThreads) cpfsk_table_kernel(float2* tab, u64 L, uint32_t n_sym, PhasorArgs p, u64 sample0)
{
    const u64 total = (u64)n_sym * L;
    for (u64 g = (u64)blockIdx.x * kThreads + threadIdx.x; g < total; g += (u64)gridDim.x * kThreads) {
        const u64 v = g / L, n = g % L;
        const float inner = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, (float)(uint32_t)v), p.deviation), __ull2float_rn(sample0 + n + 1));
        float sn, cs;
        mg_sincosf_nco(inner, &sn, &cs);
        tab[g] = make_float2(__fmul_rn(p.amplitude, cs), __fmul_rn(p.amplitude, sn));
    }
}

/*
 * Per-sample part.  One thread owns U groups of VEC consecutive samples (VEC = 2: one 128-bit store per
 * group) and loops over the CTA's frames: the carrier's (cos, sin) and, for msk, the phasor's own (cos, sin)
 * depend on the sample index only and are evaluated once per thread.  bfsk / mfsk need one sincos per sample
 * and frame (their argument depends on the frame's bits); cpfsk reads its per-(symbol value, sample) table
 * when the caller built one (PhasorArgs::cp_tab), dmpsk its per-symbol (i, q) (PhasorArgs::siq).
 */
template <int KIND, int VEC>
__global__ void __launch_bounds__(kThreads) tx_phasor_kernel(const __grid_constant__ TxArgs a, const __grid_constant__ PhasorArgs p)
{
    constexpr int U = 2;
    const u64 f0 = (u64)blockIdx.y * a.frames_per_block;
    const u64 f1 = min(a.F, f0 + a.frames_per_block);
    const float w = chan_w(a.ch, f0);

    u64 n0[U];
    float cs[U][VEC], sn[U][VEC], sf[U][VEC], pc[U][VEC], ps[U][VEC];
    uint32_t ki[U][VEC], kq[U][VEC];
    bool qv[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        n0[u] = (((u64)blockIdx.x * U + u) * kThreads + threadIdx.x) * VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const u64 n = n0[u] + v;
            ki[u][v] = (uint32_t)(n / a.sps);
            qv[u][v] = n >= a.q_offset;
            kq[u][v] = qv[u][v] ? (uint32_t)((n - a.q_offset) / a.sps) : 0u;
            mg_sincosf(nco_phase(w, a.sample0 + n), &sn[u][v], &cs[u][v]);
            sf[u][v] = __ull2float_rn(a.sample0 + n + 1); /* the phasor's `s as f32` (modulator.rs:86-97) */
            pc[u][v] = ps[u][v] = 0.0f;
            if (KIND == kPhMsk) /* msk.rs:21-23: PI / 2.0 * s as f32 / samples_per_bit as f32 */
                mg_sincosf(__fdiv_rn(__fmul_rn(kPi / 2.0f, sf[u][v]), p.samples_per_bit), &ps[u][v], &pc[u][v]);
        }
    }

    for (u64 f = f0; f < f1; ++f) {
        const uint8_t* fb = a.bits + f * a.nbits;
        const float* st = p.state + f * a.nsym;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (n0[u] >= a.L) continue;
            float2 bb[VEC], out[VEC];
            uint32_t sym_prev = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (VEC == 2 && v == 1 && n0[u] + 1 >= a.L) { /* only when L is odd: VEC == 1 is used then */
                    bb[v] = out[v] = make_float2(0.0f, 0.0f);
                    continue;
                }
                float bi, bq;
                if (KIND == kPhMsk) { /* msk.rs:29-35, bits through EvenOddOffset (data.rs:102-122) */
                    const uint32_t b0 = __ldg(fb + (u64)ki[u][v] * 2) & 1u;
                    const uint32_t b1 = qv[u][v] ? (__ldg(fb + (u64)kq[u][v] * 2 + 1) & 1u) : 0u;
                    bi = __fmul_rn(__fmul_rn(p.amplitude, (float)(2 * (int)b0 - 1)), pc[u][v]);
                    bq = __fmul_rn(__fmul_rn(-p.amplitude, (float)(2 * (int)b1 - 1)), ps[u][v]);
                } else if (KIND == kPhDmpsk) { /* dmpsk.rs:35-41 via the per-symbol table */
                    const float2 t = __ldg(p.siq + f * a.nsym + ki[u][v]);
                    bi = t.x;
                    bq = t.y;
                } else {
                    const uint32_t sym = (v > 0 && ki[u][v] == ki[u][v > 0 ? v - 1 : 0]) ? sym_prev /* same symbol as the previous sample */
                                                                                       : pack_symbol(fb + (u64)ki[u][v] * p.bps, p.bps);
                    sym_prev = sym;
                    if (KIND == kPhCpfsk && p.cp_tab) {
                        const float2 t = __ldg(p.cp_tab + (u64)sym * a.L + n0[u] + v);
                        bi = t.x;
                        bq = t.y;
                    } else {
                        float inner;
                        if (KIND == kPhBfsk) /* bfsk.rs:23-29: b as f32 * deviation * s as f32 + phase */
                            inner = __fadd_rn(__fmul_rn(__fmul_rn((float)sym, p.deviation), sf[u][v]), __ldg(st + ki[u][v]));
                        else if (KIND == kPhMfsk) /* mfsk.rs:60-62 */
                            inner = __fadd_rn(__fmul_rn(__fmul_rn(mfsk_coef(p, sym), p.deviation), sf[u][v]), __ldg(st + ki[u][v]));
                        else /* cpfsk.rs:25-31: coef = 2.0 * symbol as f32 */
                            inner = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, (float)sym), p.deviation), sf[u][v]);
                        float s_, c_;
                        mg_sincosf_nco(inner, &s_, &c_);
                        bi = __fmul_rn(p.amplitude, c_);
                        bq = __fmul_rn(p.amplitude, s_);
                    }
                }
                bb[v] = make_float2(bi, bq);
                out[v] = mix_iq(bi, bq, cs[u][v], sn[u][v]);
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
        mg_sincosf_nco(inner, &sn, &cs);
        const float c_re = cs, c_im = -sn; /* carrier.conj() */
        const float m_re = __fsub_rn(__fmul_rn(x_re, c_re), __fmul_rn(x_im, c_im)); /* num-0.1.35 Complex Mul */
        const float m_im = __fadd_rn(__fmul_rn(x_re, c_im), __fmul_rn(x_im, c_re));
        const float err = mg_atan2f(m_im, m_re);                                    /* arg() = im.atan2(re) */
        po = __fadd_rn(po, __fmul_rn(CHANGE, err));
    }
    a.po[f] = po;
}

} /* namespace mg */
