/*
 * modem_oracle.c -- CPU ORACLE (test infrastructure; see modem_oracle.h header).
 * Plain-C restatement of ramtej/rust-modem's per-sample path.  Every function cites
 * the reference lines it follows (paths relative to /root/reference/).
 *
 * Arithmetic rules kept from the reference: all intermediates are binary32, no FMA
 * contraction (compile with -ffp-contract=off), expression order as written in the
 * Rust source, `as f32` conversions where the reference has them, glibc sinf/cosf.
 */
#include "modem_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846264338327950288f /* std::f32::consts::PI */

/* ------------------------------------------------------------------ util.rs */
float orc_mod_trig(float x)
{
    /* util.rs:4-5 */
    const float TWO_PI = ORC_PI * 2.0f;
    return x - TWO_PI * floorf(x / TWO_PI);
}

/* ------------------------------------------------------------------ freq.rs */
float orc_ang_freq(size_t hz)
{
    /* freq.rs:20  `2.0 * PI * self.hz as f32` == (2.0*PI) * (hz as f32) */
    return 2.0f * ORC_PI * (float)hz;
}
float orc_sample_freq(size_t hz, size_t sr)
{
    /* freq.rs:25 */
    return orc_ang_freq(hz) / (float)sr;
}

/* ----------------------------------------------------------------- rates.rs */
size_t orc_samples_per_symbol(size_t br, size_t sr)
{
    /* rates.rs:16 integer division */
    return sr / br;
}

/* --------------------------------------------------------------- carrier.rs */
void orc_carrier_new(orc_carrier_t* c, size_t hz, size_t sr)
{
    /* carrier.rs:10-15 */
    c->sample_freq = orc_sample_freq(hz, sr);
    c->sample = 0;
}
static float carrier_inner(const orc_carrier_t* c, size_t s)
{
    /* carrier.rs:17-19 */
    return orc_mod_trig(c->sample_freq * (float)s);
}
float orc_carrier_next(orc_carrier_t* c)
{
    /* carrier.rs:21-26 */
    size_t sample = c->sample;
    c->sample += 1;
    return carrier_inner(c, sample);
}

/* ------------------------------------------------------------------ data.rs */
void orc_symbol_clock_new(orc_symbol_clock_t* c, size_t sps)
{
    /* data.rs:20-25 */
    c->samples_per_symbol = sps;
    c->counter = sps - 1;
}
int orc_symbol_clock_next(orc_symbol_clock_t* c)
{
    /* data.rs:27-32 */
    c->counter += 1;
    c->counter %= c->samples_per_symbol;
    return c->counter == 0;
}

void orc_bits_new(orc_source_t* s, const uint8_t* bits, size_t nbits, size_t sps, size_t bps)
{
    /* data.rs:43-52 */
    memset(s, 0, sizeof *s);
    s->bits = bits;
    s->nbits = nbits;
    orc_symbol_clock_new(&s->clock, sps);
    s->bits_per_symbol = bps;
    s->idx = 0;
}
void orc_evenodd_new(orc_source_t* s, const uint8_t* bits, size_t nbits, size_t sps, size_t bps)
{
    /* data.rs:88-99; asserts bps == 2 and sps % bps == 0 are the caller's duty here */
    orc_bits_new(s, bits, nbits, sps, bps);
    s->evenodd = 1;
    orc_symbol_clock_new(&s->half_clock, sps / bps);
    s->cur[0] = 0;
    s->cur[1] = 0;
}
static const uint8_t* bits_slice(const orc_source_t* s)
{
    /* data.rs:54-63 */
    size_t start = (s->idx - 1) * s->bits_per_symbol;
    size_t end = start + s->bits_per_symbol;
    return end <= s->nbits ? s->bits + start : NULL;
}
static orc_update_t bits_next(orc_source_t* s)
{
    /* data.rs:67-78 */
    orc_update_t u;
    u.len = s->bits_per_symbol;
    if (orc_symbol_clock_next(&s->clock)) {
        s->idx += 1;
        u.bits = bits_slice(s);
        u.kind = u.bits ? ORC_CHANGED : ORC_FINISHED;
    } else {
        u.bits = bits_slice(s);
        u.kind = ORC_UNCHANGED;
    }
    return u;
}
orc_update_t orc_source_next(orc_source_t* s)
{
    if (!s->evenodd) return bits_next(s);
    /* data.rs:103-122 */
    orc_update_t in = bits_next(s);
    orc_update_t u;
    u.len = 2;
    u.bits = s->cur;
    if (in.kind == ORC_FINISHED) {
        u.kind = ORC_FINISHED;
        u.bits = NULL;
        return u;
    }
    if (in.kind == ORC_CHANGED) {
        orc_symbol_clock_next(&s->half_clock);
        s->cur[0] = in.bits[0];
        u.kind = ORC_CHANGED;
        return u;
    }
    if (orc_symbol_clock_next(&s->half_clock)) {
        s->cur[1] = in.bits[1];
        u.kind = ORC_CHANGED;
    } else {
        u.kind = ORC_UNCHANGED;
    }
    return u;
}

/* ---------------------------------------------------------- digital/util.rs */
float orc_bit_to_sign(uint8_t b)
{
    /* digital/util.rs:2  (2 * b as i8 - 1) as f32 */
    return (float)(int8_t)(2 * (int8_t)b - 1);
}
uint8_t orc_bytes_to_bits(const uint8_t* b, size_t n)
{
    /* digital/util.rs:5-11, MSB first */
    size_t len = n - 1;
    uint8_t s = 0;
    for (size_t i = 0; i < n; ++i) s |= (uint8_t)((b[i] & 1) << (len - i));
    return s;
}
size_t orc_max_symbol(size_t bits_per_symbol)
{
    /* digital/util.rs:13-15 */
    return ((size_t)1 << bits_per_symbol) - 1;
}

/* ------------------------------------------------------------- constructors */
static void phasor_zero(orc_phasor_t* p, int scheme, size_t bps)
{
    memset(p, 0, sizeof *p);
    p->scheme = scheme;
    p->bits_per_symbol = bps;
}
void orc_bask_new(orc_phasor_t* p, float a)
{
    phasor_zero(p, ORC_BASK, 1); /* bask.rs:16 */
    p->amplitude = a;
}
void orc_bpsk_new(orc_phasor_t* p, float phase, float amplitude)
{
    phasor_zero(p, ORC_BPSK, 1); /* bpsk.rs:23 */
    p->phase = phase;
    p->amplitude = amplitude;
}
void orc_qpsk_new(orc_phasor_t* p, float phase, float amplitude)
{
    phasor_zero(p, ORC_QPSK, 2); /* qpsk.rs:21 */
    p->phase_cos = cosf(phase);               /* qpsk.rs:13 */
    p->phase_sin = sinf(phase);               /* qpsk.rs:14 */
    p->amplitude = amplitude * sqrtf(0.5f);   /* qpsk.rs:15 */
}
void orc_qam_new(orc_phasor_t* p, size_t bps, float phase, float amplitude)
{
    phasor_zero(p, ORC_QAM, bps);
    size_t cs = bps / 2;                      /* qam.rs:19 */
    float ms = (float)orc_max_symbol(cs);     /* qam.rs:20 */
    p->bits_per_carrier = cs;
    p->max_symbol = ms;
    p->phase_cos = cosf(phase);
    p->phase_sin = sinf(phase);
    p->amplitude = amplitude / ms / 2.0f;     /* qam.rs:28 */
}
void orc_mpsk_new(orc_phasor_t* p, size_t bps, float phase_offset, float amplitude)
{
    phasor_zero(p, ORC_MPSK, bps);
    p->num_symbols = (float)(1 << bps);       /* mpsk.rs:17 */
    p->amplitude = amplitude;
    p->phase = phase_offset;
}
void orc_oqpsk_new(orc_phasor_t* p, float amplitude)
{
    phasor_zero(p, ORC_OQPSK, 2);
    p->amplitude = amplitude * sqrtf(0.5f);   /* oqpsk.rs:11 */
}
void orc_dcqpsk_new(orc_phasor_t* p, float amplitude)
{
    phasor_zero(p, ORC_DCQPSK, 2);
    p->amplitude = amplitude;
    p->even = 0;                              /* dcqpsk.rs:19 */
}
int orc_apsk_new(orc_phasor_t* p, float amplitude, size_t bps, const orc_ring_t* rings, size_t n)
{
    phasor_zero(p, ORC_APSK, bps);
    if (n > 8) return 0;
    /* apsk.rs:85-97 verify() */
    size_t prev = 0;
    for (size_t r = 0; r < n; ++r) {
        if (rings[r].start != prev) return 0;
        /* apsk.rs:74 Ring::new assert */
        if (!(rings[r].radius >= 0.0f && rings[r].radius <= 1.0f)) return 0;
        prev = rings[r].end;
        p->rings[r] = rings[r];
    }
    if (prev != orc_max_symbol(bps) + 1) return 0;
    p->n_rings = n;
    p->amplitude = amplitude;
    return 1;
}
void orc_bfsk_new(orc_phasor_t* p, size_t dev_hz, size_t sr, float a)
{
    phasor_zero(p, ORC_BFSK, 1);
    p->deviation = orc_sample_freq(dev_hz, sr); /* bfsk.rs:16 */
    p->amplitude = a;
    p->phase = 0.0f;
    p->prev = 0;
}
void orc_mfsk_new(orc_phasor_t* p, size_t bps, size_t dev_hz, size_t sr, float a, int increase_map)
{
    phasor_zero(p, increase_map ? ORC_MFSK_INCREASE : ORC_MFSK_DEFAULT, bps);
    p->deviation = orc_sample_freq(dev_hz, sr); /* mfsk.rs:52 */
    p->amplitude = a;
    p->phase = 0.0f;                            /* phase_offset, mfsk.rs:55 */
    p->cur_coef = 0.0f;
    p->max_symbol_i = (int)orc_max_symbol(bps); /* mfsk.rs:18 */
}
void orc_cpfsk_new(orc_phasor_t* p, size_t bps, size_t br, size_t sr, float a, size_t deviation)
{
    phasor_zero(p, ORC_CPFSK, bps);
    p->deviation = orc_sample_freq(deviation * br / 2, sr); /* cpfsk.rs:19-20 */
    p->amplitude = a;
}
void orc_msk_new(orc_phasor_t* p, float a, size_t sps)
{
    phasor_zero(p, ORC_MSK, 2);
    p->amplitude = a;
    p->samples_per_bit = sps / 2;             /* msk.rs:17 */
}
void orc_dmpsk_new(orc_phasor_t* p, size_t bps, float a, float phase, float shift)
{
    phasor_zero(p, ORC_DMPSK, bps);
    p->amplitude = a;
    p->phase = phase;
    p->shift = shift;
}

int orc_phasor_by_name(orc_phasor_t* p, const char* name, size_t br, size_t sr)
{
    /* src/bin/modulate.rs:74-95 with AMPLITUDE = 1.0 (modulate.rs:14) */
    const float A = 1.0f;
    size_t sps = orc_samples_per_symbol(br, sr);
    if (!strcmp(name, "bask")) orc_bask_new(p, A);
    else if (!strcmp(name, "bpsk")) orc_bpsk_new(p, ORC_PI / 4.0f, A);
    else if (!strcmp(name, "bfsk")) orc_bfsk_new(p, 200, sr, A);
    else if (!strcmp(name, "qpsk")) orc_qpsk_new(p, 0.0f, A);
    else if (!strcmp(name, "qam16")) orc_qam_new(p, 4, 0.0f, A);
    else if (!strcmp(name, "qam256")) orc_qam_new(p, 8, 0.0f, A);
    else if (!strcmp(name, "msk")) orc_msk_new(p, A, sps);
    else if (!strcmp(name, "mfsk")) orc_mfsk_new(p, 4, 50, sr, A, 1);
    else if (!strcmp(name, "16psk")) orc_mpsk_new(p, 4, 0.0f, A);
    else if (!strcmp(name, "oqpsk")) orc_oqpsk_new(p, A);
    else if (!strcmp(name, "dcqpsk")) orc_dcqpsk_new(p, A);
    else if (!strcmp(name, "16cpfsk")) orc_cpfsk_new(p, 4, br, sr, A, 1);
    else if (!strcmp(name, "16apsk")) {
        orc_ring_t rings[2] = {{0, 4, 0.5f, ORC_PI / 4.0f}, {4, 16, 1.0f, ORC_PI / 12.0f}};
        return orc_apsk_new(p, A, 4, rings, 2);
    } else if (!strcmp(name, "dqpsk")) orc_dmpsk_new(p, 2, A, ORC_PI / 4.0f, ORC_PI / 2.0f);
    else if (!strcmp(name, "dbpsk")) orc_dmpsk_new(p, 1, A, ORC_PI / 4.0f, ORC_PI);
    else return 0;
    return 1;
}

/* ------------------------------------------------------- per-scheme helpers */
static float qam_pos_bytes(const orc_phasor_t* p, const uint8_t* b, size_t n)
{
    /* qam.rs:32-38 */
    return 2.0f * (float)orc_bytes_to_bits(b, n) - p->max_symbol;
}
static float mpsk_inner(const orc_phasor_t* p, const uint8_t* b)
{
    /* mpsk.rs:23-29 */
    float phase = 2.0f * ORC_PI * (float)orc_bytes_to_bits(b, p->bits_per_symbol) / p->num_symbols;
    return phase + p->phase;
}
static float dcqpsk_term(const orc_phasor_t* p, uint8_t symbol)
{
    /* dcqpsk.rs:23-36 */
    const float MAP[4] = {0.0f, ORC_PI / 2.0f, 3.0f * ORC_PI / 2.0f, ORC_PI};
    return p->even ? MAP[symbol] + ORC_PI / 4.0f : MAP[symbol];
}
static void apsk_common(const orc_phasor_t* p, uint8_t symbol, float* radius, float* phase)
{
    /* apsk.rs:36-42 */
    const orc_ring_t* ring = NULL;
    for (size_t r = 0; r < p->n_rings; ++r)
        if (symbol >= p->rings[r].start && symbol < p->rings[r].end) {
            ring = &p->rings[r];
            break;
        }
    *phase = 2.0f * ORC_PI * (float)(uint8_t)(symbol - ring->start) /
                 (float)(uint8_t)(ring->end - ring->start) +
             ring->phase;
    *radius = ring->radius;
}
static float bfsk_rads(const orc_phasor_t* p, size_t s, uint8_t b)
{
    /* bfsk.rs:27-29 */
    return (float)b * p->deviation * (float)s;
}
static float mfsk_coef(const orc_phasor_t* p, uint8_t symbol)
{
    if (p->scheme == ORC_MFSK_INCREASE) return (float)(uint8_t)(2 * symbol); /* mfsk.rs:32-34 (u8 arithmetic) */
    return (float)(2 * (int)symbol - p->max_symbol_i);                        /* mfsk.rs:24-26 */
}
static float mfsk_inner(const orc_phasor_t* p, size_t s)
{
    /* mfsk.rs:60-62 */
    return p->cur_coef * p->deviation * (float)s + p->phase;
}
static float cpfsk_inner(const orc_phasor_t* p, const uint8_t* b, size_t s)
{
    /* cpfsk.rs:25-31 */
    float coef = 2.0f * (float)orc_bytes_to_bits(b, p->bits_per_symbol);
    return coef * p->deviation * (float)s;
}
static float msk_inner(const orc_phasor_t* p, size_t s)
{
    /* msk.rs:21-23 */
    return ORC_PI / 2.0f * (float)s / (float)p->samples_per_bit;
}

void orc_phasor_update(orc_phasor_t* p, size_t s, const uint8_t* b)
{
    switch (p->scheme) {
    case ORC_DCQPSK: /* dcqpsk.rs:42-44 */
        p->even = !p->even;
        break;
    case ORC_BFSK: /* bfsk.rs:43-55 */
        if (b[0] == p->prev) return;
        p->phase = orc_mod_trig(p->phase + (b[0] == 1 ? -bfsk_rads(p, s, 1) : bfsk_rads(p, s - 1, 1)));
        p->prev = b[0];
        break;
    case ORC_MFSK_DEFAULT:
    case ORC_MFSK_INCREASE: { /* mfsk.rs:68-75 */
        float next_coef = mfsk_coef(p, orc_bytes_to_bits(b, p->bits_per_symbol));
        p->phase += (p->cur_coef - next_coef) * p->deviation * (float)s;
        p->phase = orc_mod_trig(p->phase);
        p->cur_coef = next_coef;
        break;
    }
    case ORC_DMPSK: /* dmpsk.rs:29-33 */
        p->phase = orc_mod_trig(p->phase + (float)orc_bytes_to_bits(b, p->bits_per_symbol) * p->shift);
        break;
    default: /* digital/phasor.rs:4 default no-op */
        break;
    }
}

float orc_phasor_i(const orc_phasor_t* p, size_t s, const uint8_t* b)
{
    switch (p->scheme) {
    case ORC_BASK: /* bask.rs:18-20 */
        return (float)b[0] * p->amplitude;
    case ORC_BPSK: /* bpsk.rs:17-19,25-27 */
        return orc_bit_to_sign(b[0]) * p->amplitude * cosf(p->phase);
    case ORC_QPSK: /* qpsk.rs:23-28 */
        return p->amplitude * (orc_bit_to_sign(b[0]) * p->phase_cos - orc_bit_to_sign(b[1]) * p->phase_sin);
    case ORC_QAM: { /* qam.rs:44-51 */
        size_t c = p->bits_per_carrier;
        return p->amplitude * (qam_pos_bytes(p, b, c) * p->phase_cos -
                               qam_pos_bytes(p, b + c, p->bits_per_symbol - c) * p->phase_sin);
    }
    case ORC_MPSK: /* mpsk.rs:35-37 */
        return p->amplitude * cosf(mpsk_inner(p, b));
    case ORC_OQPSK: /* oqpsk.rs:19-21 */
        return orc_bit_to_sign(b[0]) * p->amplitude;
    case ORC_DCQPSK: /* dcqpsk.rs:46-48 */
        return p->amplitude * cosf(dcqpsk_term(p, orc_bytes_to_bits(b, 2)));
    case ORC_APSK: { /* apsk.rs:48-51 */
        float r, inner;
        apsk_common(p, orc_bytes_to_bits(b, p->bits_per_symbol), &r, &inner);
        return p->amplitude * r * cosf(inner);
    }
    case ORC_BFSK: /* bfsk.rs:23-25,35-37 */
        return p->amplitude * cosf(bfsk_rads(p, s, b[0]) + p->phase);
    case ORC_MFSK_DEFAULT:
    case ORC_MFSK_INCREASE: /* mfsk.rs:77-79 */
        return p->amplitude * cosf(mfsk_inner(p, s));
    case ORC_CPFSK: /* cpfsk.rs:37-39 */
        return p->amplitude * cosf(cpfsk_inner(p, b, s));
    case ORC_MSK: /* msk.rs:29-31 */
        return p->amplitude * orc_bit_to_sign(b[0]) * cosf(msk_inner(p, s));
    case ORC_DMPSK: /* dmpsk.rs:35-37 */
        return p->amplitude * cosf(p->phase);
    }
    return 0.0f;
}

float orc_phasor_q(const orc_phasor_t* p, size_t s, const uint8_t* b)
{
    switch (p->scheme) {
    case ORC_BASK: /* bask.rs:22-24 */
        return 0.0f;
    case ORC_BPSK: /* bpsk.rs:29-31 */
        return orc_bit_to_sign(b[0]) * p->amplitude * sinf(p->phase);
    case ORC_QPSK: /* qpsk.rs:30-35 */
        return p->amplitude * (orc_bit_to_sign(b[1]) * p->phase_cos + orc_bit_to_sign(b[0]) * p->phase_sin);
    case ORC_QAM: { /* qam.rs:53-60 */
        size_t c = p->bits_per_carrier;
        return p->amplitude * (qam_pos_bytes(p, b + c, p->bits_per_symbol - c) * p->phase_cos +
                               qam_pos_bytes(p, b, c) * p->phase_sin);
    }
    case ORC_MPSK: /* mpsk.rs:39-41 */
        return p->amplitude * sinf(mpsk_inner(p, b));
    case ORC_OQPSK: /* oqpsk.rs:23-25 */
        return orc_bit_to_sign(b[1]) * p->amplitude;
    case ORC_DCQPSK: /* dcqpsk.rs:50-52 */
        return p->amplitude * sinf(dcqpsk_term(p, orc_bytes_to_bits(b, 2)));
    case ORC_APSK: { /* apsk.rs:53-56 */
        float r, inner;
        apsk_common(p, orc_bytes_to_bits(b, p->bits_per_symbol), &r, &inner);
        return p->amplitude * r * sinf(inner);
    }
    case ORC_BFSK: /* bfsk.rs:39-41 */
        return p->amplitude * sinf(bfsk_rads(p, s, b[0]) + p->phase);
    case ORC_MFSK_DEFAULT:
    case ORC_MFSK_INCREASE: /* mfsk.rs:81-83 */
        return p->amplitude * sinf(mfsk_inner(p, s));
    case ORC_CPFSK: /* cpfsk.rs:41-43 */
        return p->amplitude * sinf(cpfsk_inner(p, b, s));
    case ORC_MSK: /* msk.rs:33-35 */
        return -p->amplitude * orc_bit_to_sign(b[1]) * sinf(msk_inner(p, s));
    case ORC_DMPSK: /* dmpsk.rs:39-41 */
        return p->amplitude * sinf(p->phase);
    }
    return 0.0f;
}

/* ------------------------------------------------------------------- fir.rs */
int orc_fir_new(orc_fir_t* f, const float* coefs, size_t n)
{
    /* fir.rs:10-16 */
    f->coefs = coefs;
    f->n = n;
    f->history = (float*)calloc(n ? n : 1, sizeof(float));
    f->idx = 0;
    return f->history != NULL;
}
static float fir_calc(const orc_fir_t* f)
{
    /* fir.rs:18-25: cur = min(cur - 1 (wrapping), len - 1); s + history[cur] * coef */
    size_t cur = f->idx;
    float s = 0.0f;
    for (size_t k = 0; k < f->n; ++k) {
        size_t dec = cur - 1; /* wraps to SIZE_MAX when cur == 0, as in --release Rust */
        cur = dec < f->n - 1 ? dec : f->n - 1;
        s = s + f->history[cur] * f->coefs[k];
    }
    return s;
}
float orc_fir_add(orc_fir_t* f, float sample)
{
    /* fir.rs:27-34 */
    f->history[f->idx] = sample;
    f->idx += 1;
    f->idx %= f->n;
    return fir_calc(f);
}
void orc_fir_free(orc_fir_t* f)
{
    free(f->history);
    f->history = NULL;
}

/* ------------------------------------------------------------- modulator.rs */
void orc_iq_modulate(const orc_iq_sample_t* s, float* re, float* im)
{
    /* modulator.rs:45-48 sin_cos(); :37-43 real()/imag() */
    float sn = sinf(s->carrier), cs = cosf(s->carrier);
    *re = s->i * cs - s->q * sn;
    *im = s->i * sn + s->q * cs;
}
int orc_digital_modulator_next(orc_carrier_t* c, orc_phasor_t* p, orc_source_t* src, orc_iq_sample_t* out)
{
    /* modulator.rs:85-100 */
    float phase = orc_carrier_next(c);
    orc_update_t u = orc_source_next(src);
    if (u.kind == ORC_FINISHED) return 0;
    if (u.kind == ORC_CHANGED) orc_phasor_update(p, c->sample, u.bits);
    out->carrier = phase;
    out->i = orc_phasor_i(p, c->sample, u.bits); /* digital/phasor.rs:9-11 */
    out->q = orc_phasor_q(p, c->sample, u.bits);
    return 1;
}

/* ---------------------------------------------------- pll.rs, demodulator.rs */
void orc_pll_handle(orc_pll_t* pll, float carrier_phase, float x_re, float x_im)
{
    /* pll.rs:16-22; num-0.1.35 Complex: conj = (re, -im),
     * mul = (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re), arg = im.atan2(re) */
    const float CHANGE = 0.447214f; /* pll.rs:3 */
    float inner = carrier_phase + pll->phase_offset;
    float c_re = cosf(inner), c_im = -sinf(inner);
    float m_re = x_re * c_re - x_im * c_im;
    float m_im = x_re * c_im + x_im * c_re;
    float err = atan2f(m_im, m_re);
    pll->phase_offset += CHANGE * err;
}
int orc_demod_new(orc_demod_t* d, orc_carrier_t carrier, const float* taps, size_t n)
{
    /* demodulator.rs:20-30 */
    d->carrier = carrier;
    d->pll.phase_offset = 0.0f;
    return orc_fir_new(&d->lpi, taps, n) && orc_fir_new(&d->lpq, taps, n);
}
void orc_demod_lock_step(orc_demod_t* d, float x_re, float x_im)
{
    /* demodulator.rs:34 */
    orc_pll_handle(&d->pll, orc_carrier_next(&d->carrier), x_re, x_im);
}
void orc_demod_next(orc_demod_t* d, float x_re, float* i, float* q)
{
    /* demodulator.rs:44-55 */
    float phase = orc_carrier_next(&d->carrier) + d->pll.phase_offset;
    *i = 2.0f * orc_fir_add(&d->lpi, x_re * cosf(phase));
    *q = 2.0f * orc_fir_add(&d->lpq, x_re * -sinf(phase));
}
void orc_demod_free(orc_demod_t* d)
{
    orc_fir_free(&d->lpi);
    orc_fir_free(&d->lpq);
}

/* ------------------------------------------------- src/bin/demodulate.rs taps */
const float* orc_lowpass_taps(size_t* n)
{
    /* demodulate.rs:82-147: f32 literals written with binary64 digits; the Rust
     * compiler rounds each literal to the nearest f32, as the C compiler does here. */
    static const float COEFS[64] = {
        8.6464950643449706e-05f, -0.0011227727551926443f, -0.0010137373532784653f, -0.00051892546397063074f,
        0.00065737693207229997f, 0.0019426724039296576f, 0.0023575316971358984f, 0.0011698129325984573f,
        -0.0014109570575621668f, -0.0040119731215088154f, -0.0047065995954001117f, -0.0022692944513388992f,
        0.0026579628895631122f, 0.0073998732470493874f, 0.0085194671337849165f, 0.0040456650224074651f,
        -0.0046645972566385554f, -0.012862659808170144f, -0.014703261637603555f, -0.0069572953029268525f,
        0.00800563700908981f, 0.022172065878291854f, 0.025574286331781385f, 0.012291851983914071f,
        -0.014450589851381347f, -0.041421606566596714f, -0.05018918856526014f, -0.025933101216317672f,
        0.03394517722329659f, 0.11612232604813434f, 0.19513123601730936f, 0.24347923270043995f,
        0.24347923270043995f, 0.19513123601730936f, 0.11612232604813434f, 0.03394517722329659f,
        -0.025933101216317672f, -0.05018918856526014f, -0.041421606566596714f, -0.014450589851381347f,
        0.012291851983914071f, 0.025574286331781385f, 0.022172065878291854f, 0.00800563700908981f,
        -0.0069572953029268525f, -0.014703261637603555f, -0.012862659808170144f, -0.0046645972566385554f,
        0.0040456650224074651f, 0.0085194671337849165f, 0.0073998732470493874f, 0.0026579628895631122f,
        -0.0022692944513388992f, -0.0047065995954001117f, -0.0040119731215088154f, -0.0014109570575621668f,
        0.0011698129325984573f, 0.0023575316971358984f, 0.0019426724039296576f, 0.00065737693207229997f,
        -0.00051892546397063074f, -0.0010137373532784653f, -0.0011227727551926443f, 8.6464950643449706e-05f};
    *n = 64;
    return COEFS;
}
const float* orc_hilbert_taps(size_t* n)
{
    /* demodulate.rs:48-72 */
    static const float COEFS[23] = {-0.007576f, -2.803e-16f, -0.019824f, 3.7096e-16f, -0.044089f, 1.3201e-16f,
                                    -0.089244f, -3.2694e-16f, -0.18728f, -1.6739e-16f, -0.62794f, 0.0f,
                                    0.62794f, 1.6739e-16f, 0.18728f, 3.2694e-16f, 0.089244f, -1.3201e-16f,
                                    0.044089f, -3.7096e-16f, 0.019824f, 2.803e-16f, 0.007576f};
    *n = 23;
    return COEFS;
}

/* ================================ extensions ================================== */

void orc_rrc_taps(float* out, size_t span, size_t sps, double beta)
{
    size_t n = span * sps + 1;
    double* h = (double*)malloc(n * sizeof(double));
    const double pi = 3.14159265358979323846;
    double energy = 0.0;
    for (size_t k = 0; k < n; ++k) {
        double t = ((double)k - (double)(n - 1) / 2.0) / (double)sps; /* in symbols */
        double v;
        if (fabs(t) < 1e-12) {
            v = 1.0 - beta + 4.0 * beta / pi;
        } else if (beta > 0.0 && fabs(fabs(t) - 1.0 / (4.0 * beta)) < 1e-9) {
            v = (beta / sqrt(2.0)) *
                ((1.0 + 2.0 / pi) * sin(pi / (4.0 * beta)) + (1.0 - 2.0 / pi) * cos(pi / (4.0 * beta)));
        } else {
            double a = 4.0 * beta * t;
            v = (sin(pi * t * (1.0 - beta)) + a * cos(pi * t * (1.0 + beta))) / (pi * t * (1.0 - a * a));
        }
        h[k] = v;
        energy += v * v;
    }
    double g = 1.0 / sqrt(energy);
    for (size_t k = 0; k < n; ++k) out[k] = (float)(h[k] * g);
    free(h);
}

static void mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo)
{
    uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(0xD2511F53u, c[0], &hi0, &lo0);
        mulhilo32(0xCD9E8D57u, c[2], &hi1, &lo1);
        uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* the normal pair of modem_oracle.h "AWGN": every operation is one correctly rounded binary32 operation */
void orc_box_muller(uint32_t r0, uint32_t r1, float* n0, float* n1)
{
    static const float L[7] = {-0x1.00001cp-1f, 0x1.555802p-2f, -0x1.ffa938p-3f, 0x1.97ecccp-3f, -0x1.5e404cp-3f, 0x1.495358p-3f, -0x1.ab64d2p-4f};
    static const float S[3] = {-0x1.555552p-3f, 0x1.110c2ap-7f, -0x1.9aca02p-13f};
    static const float C[4] = {-0x1p-1f, 0x1.55554cp-5f, -0x1.6c0e0cp-10f, 0x1.9a6fd8p-16f};
    const float LN2 = 0x1.62e43p-1f, ANG = 0x1.921fb6p-22f;
    /* radius */
    float u = ((float)(r0 >> 9) + 0.5f) * 0x1p-23f;
    uint32_t ub;
    memcpy(&ub, &u, 4);
    uint32_t ix = ub - 0x3f3504f3u;
    int32_t e = (int32_t)ix >> 23;
    uint32_t mb = (ix & 0x007fffffu) + 0x3f3504f3u;
    float m;
    memcpy(&m, &mb, 4);
    float f = m - 1.0f;
    float q = L[6];
    for (int k = 5; k >= 0; --k) q = fmaf(q, f, L[k]);
    float lnm = f * fmaf(f, q, 1.0f);
    float lnu = fmaf((float)e, LN2, lnm);
    float rad = sqrtf(-2.0f * lnu);
    /* angle */
    uint32_t j = r1 >> 8, oct = j >> 21, k = j & 0x1fffffu;
    if (oct & 1u) k = 0x200000u - k;
    float x = (float)k * ANG;
    float z = x * x;
    float sn = fmaf(x * z, fmaf(fmaf(S[2], z, S[1]), z, S[0]), x);
    float cs = fmaf(z, fmaf(fmaf(fmaf(C[3], z, C[2]), z, C[1]), z, C[0]), 1.0f);
    if ((oct + 1u) & 2u) { /* octants 1, 2, 5, 6 */
        float t = sn;
        sn = cs;
        cs = t;
    }
    if ((oct + 2u) & 4u) cs = -cs; /* octants 2..5 */
    if (oct & 4u) sn = -sn;        /* octants 4..7 */
    *n0 = rad * cs;
    *n1 = rad * sn;
}
void orc_awgn_sample(uint64_t seed, uint64_t frame, uint64_t n, float sigma, float* re, float* im)
{
    uint64_t quad = n >> 2;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    float nz[2];
    for (uint32_t rail = 0; rail < 2; ++rail) {
        uint32_t ctr[4] = {(uint32_t)quad, (uint32_t)(quad >> 32) | (rail << 31), (uint32_t)frame, (uint32_t)(frame >> 32)};
        uint32_t r[4];
        orc_philox4x32_10(ctr, key, r);
        float a0, a1;
        if (n & 2) orc_box_muller(r[2], r[3], &a0, &a1);
        else orc_box_muller(r[0], r[1], &a0, &a1);
        nz[rail] = (n & 1) ? a1 : a0;
    }
    *re = *re + sigma * nz[0];
    *im = *im + sigma * nz[1];
}
void orc_awgn(float* buf, size_t F, size_t L, float sigma, uint64_t seed, uint64_t frame0)
{
    for (size_t f = 0; f < F; ++f)
        for (size_t n = 0; n < L; ++n) {
            float* s = buf + 2 * (f * L + n);
            orc_awgn_sample(seed, frame0 + f, n, sigma, &s[0], &s[1]);
        }
}
void orc_random_bits(uint8_t* bits, size_t F, size_t nbits, uint64_t seed, uint64_t frame0)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x62697473u};
    for (size_t f = 0; f < F; ++f) {
        uint64_t g = frame0 + f;
        for (size_t b = 0; b * 128 < nbits; ++b) {
            uint32_t ctr[4] = {(uint32_t)b, (uint32_t)((uint64_t)b >> 32), (uint32_t)g, (uint32_t)(g >> 32)};
            uint32_t r[4];
            orc_philox4x32_10(ctr, key, r);
            for (size_t j = b * 128; j < nbits && j < (b + 1) * 128; ++j) bits[f * nbits + j] = (uint8_t)((r[(j % 128) / 32] >> (j % 32)) & 1u);
        }
    }
}

void orc_unpack_bits(const uint8_t* packed, uint8_t* bits, size_t F, size_t nbits)
{
    const size_t pb = (nbits + 7) / 8;
    for (size_t f = 0; f < F; ++f)
        for (size_t j = 0; j < nbits; ++j) bits[f * nbits + j] = (uint8_t)((packed[f * pb + j / 8] >> (7 - j % 8)) & 1u);
}
void orc_pack_bits(const uint8_t* bits, uint8_t* packed, size_t F, size_t nbits)
{
    const size_t pb = (nbits + 7) / 8;
    for (size_t f = 0; f < F; ++f) {
        for (size_t b = 0; b < pb; ++b) packed[f * pb + b] = 0;
        for (size_t j = 0; j < nbits; ++j) packed[f * pb + j / 8] |= (uint8_t)((bits[f * nbits + j] & 1u) << (7 - j % 8));
    }
}

float orc_sigma_for_ebn0(const float* const_iq, size_t n_points, size_t bps, float slicer_gain,
                         float rx_gain, const float* rx_taps, size_t n_rx, double ebn0_db)
{
    double es = 0.0, eh = 0.0;
    for (size_t i = 0; i < n_points; ++i)
        es += (double)const_iq[2 * i] * const_iq[2 * i] + (double)const_iq[2 * i + 1] * const_iq[2 * i + 1];
    es /= (double)n_points;
    for (size_t i = 0; i < n_rx; ++i) eh += (double)rx_taps[i] * rx_taps[i];
    double n0 = es / ((double)bps * pow(10.0, ebn0_db / 10.0));
    return (float)((double)slicer_gain * sqrt(n0 / eh) / (double)rx_gain);
}

/* ------------------------------------------------------------ batch drivers */
static int path_phasor(const orc_path_t* p, orc_phasor_t* ph)
{
    return orc_phasor_by_name(ph, p->scheme, p->baud_rate, p->sample_rate);
}
static int path_evenodd(const orc_path_t* p)
{
    /* modulate.rs:101-107 */
    return !strcmp(p->scheme, "msk") || !strcmp(p->scheme, "oqpsk");
}
size_t orc_bits_per_symbol(const orc_path_t* p)
{
    orc_phasor_t ph;
    if (!path_phasor(p, &ph)) return 0;
    return ph.bits_per_symbol;
}
size_t orc_frame_samples(const orc_path_t* p, size_t nbits)
{
    size_t bps = orc_bits_per_symbol(p);
    return bps ? (nbits / bps) * orc_samples_per_symbol(p->baud_rate, p->sample_rate) : 0;
}
static size_t path_q_offset(const orc_path_t* p)
{
    return path_evenodd(p) ? orc_samples_per_symbol(p->baud_rate, p->sample_rate) / 2 : 0;
}
size_t orc_decided_symbols(const orc_path_t* p, size_t L)
{
    size_t sps = orc_samples_per_symbol(p->baud_rate, p->sample_rate);
    size_t last = p->decision_delay + path_q_offset(p);
    if (L <= last) return 0;
    return (L - 1 - last) / sps + 1;
}
size_t orc_constellation(const orc_path_t* p, float* out_iq, size_t cap_points)
{
    orc_phasor_t ph;
    if (!path_phasor(p, &ph)) return 0;
    size_t bps = ph.bits_per_symbol, np = (size_t)1 << bps;
    size_t n_tables = ph.scheme == ORC_DCQPSK ? 2 : 1;
    if (n_tables * np > cap_points) return 0;
    uint8_t b[8];
    for (size_t t = 0; t < n_tables; ++t) {
        for (size_t sym = 0; sym < np; ++sym) {
            for (size_t j = 0; j < bps; ++j) b[j] = (uint8_t)((sym >> (bps - 1 - j)) & 1);
            orc_phasor_t q;
            path_phasor(p, &q);
            for (size_t u = 0; u <= t; ++u) orc_phasor_update(&q, 1, b);
            out_iq[2 * (t * np + sym)] = orc_phasor_i(&q, 1, b);
            out_iq[2 * (t * np + sym) + 1] = orc_phasor_q(&q, 1, b);
        }
    }
    return n_tables;
}

static int modulate_frame(const orc_path_t* p, const uint8_t* bits, size_t nbits, float* tx, float* iq)
{
    orc_phasor_t ph;
    if (!path_phasor(p, &ph)) return -1;
    size_t sps = orc_samples_per_symbol(p->baud_rate, p->sample_rate);
    size_t q_off = path_q_offset(p);
    orc_carrier_t carrier;
    orc_carrier_new(&carrier, p->carrier_hz, p->sample_rate);
    carrier.sample = p->sample0;
    orc_source_t src;
    if (path_evenodd(p)) orc_evenodd_new(&src, bits, nbits, sps, ph.bits_per_symbol);
    else orc_bits_new(&src, bits, nbits, sps, ph.bits_per_symbol);
    orc_fir_t fi, fq;
    int shaped = p->n_tx_taps > 0;
    if (shaped && !(orc_fir_new(&fi, p->tx_taps, p->n_tx_taps) && orc_fir_new(&fq, p->tx_taps, p->n_tx_taps)))
        return -2;
    orc_iq_sample_t s;
    size_t n = 0;
    while (orc_digital_modulator_next(&carrier, &ph, &src, &s)) {
        if (shaped) {
            /* extension 1: FIR (fir.rs semantics) over the zero-stuffed symbol train */
            s.i = orc_fir_add(&fi, n % sps == 0 ? s.i : 0.0f);
            s.q = orc_fir_add(&fq, n % sps == q_off ? s.q : 0.0f);
        }
        if (iq) {
            iq[2 * n] = s.i;
            iq[2 * n + 1] = s.q;
        }
        if (tx) orc_iq_modulate(&s, &tx[2 * n], &tx[2 * n + 1]);
        ++n;
    }
    if (shaped) {
        orc_fir_free(&fi);
        orc_fir_free(&fq);
    }
    return 0;
}

int orc_modulate(const orc_path_t* p, const uint8_t* bits, size_t F, size_t nbits, float* tx, float* iq)
{
    size_t L = orc_frame_samples(p, nbits);
    for (size_t f = 0; f < F; ++f) {
        int rc = modulate_frame(p, bits + f * nbits, nbits, tx ? tx + 2 * f * L : NULL, iq ? iq + 2 * f * L : NULL);
        if (rc) return rc;
    }
    return 0;
}

/* extension 4: nearest point of the gain-scaled constellation, ties -> lowest index */
static uint8_t slice_point(const float* table, size_t np, float g, float I, float Q)
{
    size_t best = 0;
    float best_d = 0.0f;
    for (size_t s = 0; s < np; ++s) {
        float di = I - g * table[2 * s];
        float dq = Q - g * table[2 * s + 1];
        float d = di * di + dq * dq;
        if (s == 0 || d < best_d) {
            best_d = d;
            best = s;
        }
    }
    return (uint8_t)best;
}

static int demodulate_frame(const orc_path_t* p, const float* rx, size_t L, float* filt, uint8_t* sym,
                            uint8_t* bits_out, const float* table, size_t n_tables)
{
    size_t sps = orc_samples_per_symbol(p->baud_rate, p->sample_rate);
    size_t bps = orc_bits_per_symbol(p), np = (size_t)1 << bps;
    size_t q_off = path_q_offset(p);
    size_t K = orc_decided_symbols(p, L);
    orc_carrier_t carrier;
    orc_carrier_new(&carrier, p->carrier_hz, p->sample_rate);
    carrier.sample = p->sample0;
    orc_demod_t d;
    if (!orc_demod_new(&d, carrier, p->rx_taps, p->n_rx_taps)) return -2;
    d.pll.phase_offset = p->phase_offset;
    float* rail_i = NULL;
    int want_dec = (sym || bits_out) && K > 0;
    if (want_dec) rail_i = (float*)malloc(K * sizeof(float));
    for (size_t n = 0; n < L; ++n) {
        float I, Q;
        orc_demod_next(&d, rx[2 * n], &I, &Q);
        if (filt) {
            filt[2 * n] = I;
            filt[2 * n + 1] = Q;
        }
        if (want_dec) {
            /* extension 3: symbol k sliced at k*sps + decision_delay (Q rail q_off later) */
            if (n >= p->decision_delay && (n - p->decision_delay) % sps == 0) {
                size_t k = (n - p->decision_delay) / sps;
                if (k < K) rail_i[k] = I;
            }
            if (n >= p->decision_delay + q_off && (n - p->decision_delay - q_off) % sps == 0) {
                size_t k = (n - p->decision_delay - q_off) / sps;
                if (k < K) {
                    uint8_t s = slice_point(table + 2 * np * (k % n_tables), np, p->slicer_gain, rail_i[k], Q);
                    if (sym) sym[k] = s;
                    if (bits_out)
                        for (size_t j = 0; j < bps; ++j) bits_out[k * bps + j] = (uint8_t)((s >> (bps - 1 - j)) & 1);
                }
            }
        }
    }
    free(rail_i);
    orc_demod_free(&d);
    return 0;
}

int orc_demodulate(const orc_path_t* p, const float* rx, size_t F, size_t L, float* filt, uint8_t* sym,
                   uint8_t* bits_out)
{
    float table[2 * 512];
    size_t n_tables = orc_constellation(p, table, 512);
    size_t bps = orc_bits_per_symbol(p);
    if (!n_tables) return -1;
    size_t K = orc_decided_symbols(p, L);
    for (size_t f = 0; f < F; ++f) {
        int rc = demodulate_frame(p, rx + 2 * f * L, L, filt ? filt + 2 * f * L : NULL, sym ? sym + f * K : NULL,
                                  bits_out ? bits_out + f * K * bps : NULL, table, n_tables);
        if (rc) return rc;
    }
    return 0;
}

typedef struct {
    const orc_path_t* p;
    const uint8_t* bits;
    size_t f0, f1, nbits, L, K, bps;
    float sigma;
    uint64_t seed, frame0;
    uint8_t *sym, *bits_out;
    const float* table;
    size_t n_tables;
    uint64_t errors, compared;
    int rc;
} loop_job_t;

static void* loopback_worker(void* arg)
{
    loop_job_t* j = (loop_job_t*)arg;
    float* tx = (float*)malloc(2 * j->L * sizeof(float) + 8);
    uint8_t* dec = (uint8_t*)malloc(j->K * j->bps + 1);
    j->rc = 0;
    for (size_t f = j->f0; f < j->f1 && !j->rc; ++f) {
        const uint8_t* b = j->bits + f * j->nbits;
        j->rc = modulate_frame(j->p, b, j->nbits, tx, NULL);
        if (j->rc) break;
        if (j->sigma > 0.0f)
            for (size_t n = 0; n < j->L; ++n)
                orc_awgn_sample(j->seed, j->frame0 + f, n, j->sigma, &tx[2 * n], &tx[2 * n + 1]);
        uint8_t* bo = j->bits_out ? j->bits_out + f * j->K * j->bps : dec;
        j->rc = demodulate_frame(j->p, tx, j->L, NULL, j->sym ? j->sym + f * j->K : NULL, bo, j->table, j->n_tables);
        for (size_t i = 0; i < j->K * j->bps; ++i) j->errors += bo[i] != b[i];
        j->compared += j->K * j->bps;
    }
    free(tx);
    free(dec);
    return NULL;
}

int orc_loopback(const orc_path_t* p, const uint8_t* bits, size_t F, size_t nbits, float sigma, uint64_t seed,
                 uint64_t frame0, int threads, uint8_t* sym, uint8_t* bits_out, uint64_t counters[2])
{
    float table[2 * 512];
    size_t n_tables = orc_constellation(p, table, 512);
    if (!n_tables) return -1;
    size_t L = orc_frame_samples(p, nbits);
    size_t K = orc_decided_symbols(p, L);
    size_t bps = orc_bits_per_symbol(p);
    if (threads < 1) threads = 1;
    if ((size_t)threads > F) threads = F ? (int)F : 1;
    loop_job_t* jobs = (loop_job_t*)calloc((size_t)threads, sizeof *jobs);
    pthread_t* tids = (pthread_t*)calloc((size_t)threads, sizeof *tids);
    for (int t = 0; t < threads; ++t) {
        loop_job_t* j = &jobs[t];
        j->p = p; j->bits = bits; j->nbits = nbits; j->L = L; j->K = K; j->bps = bps;
        j->f0 = F * (size_t)t / (size_t)threads;
        j->f1 = F * (size_t)(t + 1) / (size_t)threads;
        j->sigma = sigma; j->seed = seed; j->frame0 = frame0;
        j->sym = sym; j->bits_out = bits_out; j->table = table; j->n_tables = n_tables;
        if (threads == 1) loopback_worker(j);
        else pthread_create(&tids[t], NULL, loopback_worker, j);
    }
    int rc = 0;
    for (int t = 0; t < threads; ++t) {
        if (threads > 1) pthread_join(tids[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        counters[0] += jobs[t].errors;
        counters[1] += jobs[t].compared;
    }
    free(jobs);
    free(tids);
    return rc;
}

/* ============================ src/bin sample paths ============================ */
int orc_preamble(const orc_path_t* p, size_t F, size_t n, float amplitude, float* tx)
{
    for (size_t f = 0; f < F; ++f) {
        orc_carrier_t carrier;
        orc_carrier_new(&carrier, p->carrier_hz, p->sample_rate);
        carrier.sample = p->sample0;
        for (size_t j = 0; j < n; ++j) {
            /* modulator.rs:54-61: phase = carrier.next(); Raw::next => (amplitude, 0.0) (phasor.rs:17-23) */
            orc_iq_sample_t s = {orc_carrier_next(&carrier), amplitude, 0.0f};
            orc_iq_modulate(&s, &tx[2 * (f * n + j)], &tx[2 * (f * n + j) + 1]);
        }
    }
    return 0;
}

int orc_modulate_real(const orc_path_t* p, const uint8_t* bits, size_t F, size_t nbits, size_t preamble,
                      float preamble_amplitude, float* out)
{
    orc_phasor_t ph0;
    if (!path_phasor(p, &ph0)) return -1;
    size_t sps = orc_samples_per_symbol(p->baud_rate, p->sample_rate);
    size_t L = orc_frame_samples(p, nbits);
    for (size_t f = 0; f < F; ++f) {
        float* o = out + f * (preamble + L);
        orc_carrier_t carrier; /* ONE carrier for tone and data (modulate.rs:71,120,128) */
        orc_carrier_new(&carrier, p->carrier_hz, p->sample_rate);
        carrier.sample = p->sample0;
        float im;
        for (size_t j = 0; j < preamble; ++j) { /* modulate.rs:120-125 */
            orc_iq_sample_t s = {orc_carrier_next(&carrier), preamble_amplitude, 0.0f};
            orc_iq_modulate(&s, &o[j], &im);
        }
        orc_phasor_t ph = ph0;
        orc_source_t src;
        if (path_evenodd(p)) orc_evenodd_new(&src, bits + f * nbits, nbits, sps, ph.bits_per_symbol);
        else orc_bits_new(&src, bits + f * nbits, nbits, sps, ph.bits_per_symbol);
        orc_iq_sample_t s;
        size_t n = 0;
        while (orc_digital_modulator_next(&carrier, &ph, &src, &s)) { /* modulate.rs:128-133 */
            if (n >= L) return -4;
            orc_iq_modulate(&s, &o[preamble + n], &im);
            ++n;
        }
        if (n != L) return -4;
    }
    return 0;
}

int orc_demodulate_real(const orc_path_t* p, const float* x, const float* analytic_im, size_t F, size_t L, size_t lock,
                        const float* hilbert, size_t n_hilbert, float* po_out, float* filt, uint8_t* sym, uint8_t* bits_out)
{
    if (L < lock) return -3;
    float table[2 * 512];
    size_t n_tables = orc_constellation(p, table, 512);
    if (!n_tables) return -1;
    size_t sps = orc_samples_per_symbol(p->baud_rate, p->sample_rate);
    size_t bps = orc_bits_per_symbol(p), np = (size_t)1 << bps;
    size_t q_off = path_q_offset(p);
    size_t Lr = L - lock;
    size_t K = orc_decided_symbols(p, Lr);
    if (!hilbert) hilbert = orc_hilbert_taps(&n_hilbert);
    float* rail_i = (float*)malloc((K ? K : 1) * sizeof(float));
    for (size_t f = 0; f < F; ++f) {
        const float* xf = x + f * L;
        orc_fir_t hfir; /* demodulate.rs:31 */
        if (!orc_fir_new(&hfir, hilbert, n_hilbert)) return -2;
        orc_carrier_t carrier;
        orc_carrier_new(&carrier, p->carrier_hz, p->sample_rate);
        carrier.sample = p->sample0;
        orc_demod_t d;
        if (!orc_demod_new(&d, carrier, p->rx_taps, p->n_rx_taps)) return -2;
        if (!lock) d.pll.phase_offset = p->phase_offset;
        size_t t = 0;
        for (; t < lock; ++t) { /* demodulator.rs:32-36 */
            float im = analytic_im ? analytic_im[f * L + t] : orc_fir_add(&hfir, xf[t]); /* demodulate.rs:32-34 */
            orc_demod_lock_step(&d, xf[t], im);
        }
        if (po_out) po_out[f] = d.pll.phase_offset;
        for (size_t n = 0; n < Lr; ++n) { /* demodulate.rs:41-43 over demodulator.rs:44-55 */
            float I, Q;
            orc_demod_next(&d, xf[lock + n], &I, &Q);
            if (filt) {
                filt[2 * (f * Lr + n)] = I;
                filt[2 * (f * Lr + n) + 1] = Q;
            }
            if ((sym || bits_out) && K) {
                if (n >= p->decision_delay && (n - p->decision_delay) % sps == 0) {
                    size_t k = (n - p->decision_delay) / sps;
                    if (k < K) rail_i[k] = I;
                }
                if (n >= p->decision_delay + q_off && (n - p->decision_delay - q_off) % sps == 0) {
                    size_t k = (n - p->decision_delay - q_off) / sps;
                    if (k < K) {
                        uint8_t s = slice_point(table + 2 * np * (k % n_tables), np, p->slicer_gain, rail_i[k], Q);
                        if (sym) sym[f * K + k] = s;
                        if (bits_out)
                            for (size_t j = 0; j < bps; ++j) bits_out[(f * K + k) * bps + j] = (uint8_t)((s >> (bps - 1 - j)) & 1);
                    }
                }
            }
        }
        orc_demod_free(&d);
        orc_fir_free(&hfir);
    }
    free(rail_i);
    return 0;
}
