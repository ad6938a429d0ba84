/*
 * host_tables.cpp -- host-side constants of the path, computed the reference's way.
 *
 * These are the values a caller of the reference obtains from Freq / Rates and from the
 * digital::* constructors + i()/q() (paths relative to /root/reference/src/modem/).  They
 * are evaluated once per configuration on the host in binary32 with glibc sinf/cosf (what
 * Rust's f32::sin/cos lower to) and handed to the kernels as a constellation table.
 * Compile with -ffp-contract=off: the reference never fuses a*b+c.
 */
#include <math.h>
#include <string.h>

#include "../../include/modem_gpu.h"

namespace {
const float PI32 = 3.14159265358979323846264338327950288f; /* std::f32::consts::PI */

inline float bit_to_sign(uint32_t b) { return (float)(2 * (int)b - 1); } /* digital/util.rs:1-3 */
inline uint32_t bit(uint32_t sym, uint32_t bps, uint32_t j) { return (sym >> (bps - 1 - j)) & 1u; } /* util.rs:5-11 */
} // namespace

extern "C" {

float modem_sample_freq(size_t hz, size_t sr)
{
    /* freq.rs:20,25 */
    float ang = 2.0f * PI32 * (float)hz;
    return ang / (float)sr;
}

size_t modem_samples_per_symbol(size_t baud_rate, size_t sample_rate)
{
    return baud_rate ? sample_rate / baud_rate : 0; /* rates.rs:16 */
}

int modem_const_bask(float amplitude, float* out)
{
    for (uint32_t s = 0; s < 2; ++s) {
        out[2 * s] = (float)s * amplitude; /* bask.rs:19 */
        out[2 * s + 1] = 0.0f;             /* bask.rs:23 */
    }
    return 1;
}

int modem_const_bpsk(float phase, float amplitude, float* out)
{
    for (uint32_t s = 0; s < 2; ++s) {
        float common = bit_to_sign(s) * amplitude; /* bpsk.rs:18 */
        out[2 * s] = common * cosf(phase);         /* bpsk.rs:26 */
        out[2 * s + 1] = common * sinf(phase);     /* bpsk.rs:30 */
    }
    return 1;
}

int modem_const_qpsk(float phase, float amplitude, float* out)
{
    const float pc = cosf(phase), ps = sinf(phase); /* qpsk.rs:13-14 */
    const float a = amplitude * sqrtf(0.5f);        /* qpsk.rs:15 */
    for (uint32_t s = 0; s < 4; ++s) {
        float s0 = bit_to_sign(bit(s, 2, 0)), s1 = bit_to_sign(bit(s, 2, 1));
        out[2 * s] = a * (s0 * pc - s1 * ps);     /* qpsk.rs:24-27 */
        out[2 * s + 1] = a * (s1 * pc + s0 * ps); /* qpsk.rs:31-34 */
    }
    return 2;
}

int modem_const_qam(uint32_t bps, float phase, float amplitude, float* out)
{
    if (bps < 2 || bps > 8) return MODEM_ERR_INVALID; /* qam.rs:17 assert */
    const uint32_t cs = bps / 2;                      /* qam.rs:19 */
    const float ms = (float)((1u << cs) - 1u);        /* qam.rs:20 */
    const float pc = cosf(phase), ps = sinf(phase);
    const float a = amplitude / ms / 2.0f;            /* qam.rs:28 */
    const uint32_t lsb_bits = bps - cs;
    for (uint32_t s = 0; s < (1u << bps); ++s) {
        uint32_t msb = s >> lsb_bits, lsb = s & ((1u << lsb_bits) - 1u); /* split_at(bits_per_carrier) */
        float pm = 2.0f * (float)msb - ms, pl = 2.0f * (float)lsb - ms;  /* qam.rs:32-34 */
        out[2 * s] = a * (pm * pc - pl * ps);                            /* qam.rs:47-50 */
        out[2 * s + 1] = a * (pl * pc + pm * ps);                        /* qam.rs:56-59 */
    }
    return (int)bps;
}

int modem_const_mpsk(uint32_t bps, float phase_offset, float amplitude, float* out)
{
    if (bps < 1 || bps > 8) return MODEM_ERR_INVALID;
    const float num = (float)(1u << bps); /* mpsk.rs:17 */
    for (uint32_t s = 0; s < (1u << bps); ++s) {
        float inner = 2.0f * PI32 * (float)s / num + phase_offset; /* mpsk.rs:24,28 */
        out[2 * s] = amplitude * cosf(inner);
        out[2 * s + 1] = amplitude * sinf(inner);
    }
    return (int)bps;
}

int modem_const_oqpsk(float amplitude, float* out)
{
    const float a = amplitude * sqrtf(0.5f); /* oqpsk.rs:11 */
    for (uint32_t s = 0; s < 4; ++s) {
        out[2 * s] = bit_to_sign(bit(s, 2, 0)) * a;     /* oqpsk.rs:20 */
        out[2 * s + 1] = bit_to_sign(bit(s, 2, 1)) * a; /* oqpsk.rs:24 */
    }
    return 2;
}

int modem_const_dcqpsk(float amplitude, float* out)
{
    /* dcqpsk.rs:24-29; `even` is toggled by update() before the first symbol is evaluated
     * (modulator.rs:90-91), so table 0 (symbol 0, 2, ...) is the +pi/4 constellation. */
    const float MAP[4] = {0.0f, PI32 / 2.0f, 3.0f * PI32 / 2.0f, PI32};
    for (uint32_t t = 0; t < 2; ++t)
        for (uint32_t s = 0; s < 4; ++s) {
            float term = t == 0 ? MAP[s] + PI32 / 4.0f : MAP[s];
            out[2 * (4 * t + s)] = amplitude * cosf(term);
            out[2 * (4 * t + s) + 1] = amplitude * sinf(term);
        }
    return 2;
}

int modem_const_apsk(float amplitude, uint32_t bps, const modem_ring_t* rings, size_t n_rings, float* out)
{
    if (bps < 1 || bps > 8 || !rings) return MODEM_ERR_INVALID;
    /* apsk.rs:85-97 verify(), apsk.rs:74 radius assert */
    uint32_t prev = 0;
    for (size_t r = 0; r < n_rings; ++r) {
        if (rings[r].start != prev) return MODEM_ERR_INVALID;
        if (!(rings[r].radius >= 0.0f && rings[r].radius <= 1.0f)) return MODEM_ERR_INVALID;
        prev = rings[r].end;
    }
    if (prev != (1u << bps)) return MODEM_ERR_INVALID;
    for (uint32_t s = 0; s < (1u << bps); ++s) {
        const modem_ring_t* ring = nullptr;
        for (size_t r = 0; r < n_rings; ++r)
            if (s >= rings[r].start && s < rings[r].end) {
                ring = &rings[r];
                break;
            }
        float phase = 2.0f * PI32 * (float)(s - ring->start) / (float)(ring->end - ring->start) + ring->phase; /* apsk.rs:38-39 */
        out[2 * s] = amplitude * ring->radius * cosf(phase);     /* apsk.rs:50 */
        out[2 * s + 1] = amplitude * ring->radius * sinf(phase); /* apsk.rs:55 */
    }
    return (int)bps;
}

int modem_const_by_name(const char* name, float* out, uint32_t* n_tables, uint32_t* evenodd)
{
    /* src/bin/modulate.rs:74-95, AMPLITUDE = 1.0 (modulate.rs:14) */
    const float A = 1.0f;
    if (!name || !out) return MODEM_ERR_INVALID;
    if (n_tables) *n_tables = 1;
    if (evenodd) *evenodd = 0;
    if (!strcmp(name, "bask")) return modem_const_bask(A, out);
    if (!strcmp(name, "bpsk")) return modem_const_bpsk(PI32 / 4.0f, A, out);
    if (!strcmp(name, "qpsk")) return modem_const_qpsk(0.0f, A, out);
    if (!strcmp(name, "qam16")) return modem_const_qam(4, 0.0f, A, out);
    if (!strcmp(name, "qam256")) return modem_const_qam(8, 0.0f, A, out);
    if (!strcmp(name, "16psk")) return modem_const_mpsk(4, 0.0f, A, out);
    if (!strcmp(name, "oqpsk")) {
        if (evenodd) *evenodd = 1; /* modulate.rs:103 */
        return modem_const_oqpsk(A, out);
    }
    if (!strcmp(name, "dcqpsk")) {
        if (n_tables) *n_tables = 2;
        return modem_const_dcqpsk(A, out);
    }
    if (!strcmp(name, "16apsk")) {
        const modem_ring_t rings[2] = {{0, 4, 0.5f, PI32 / 4.0f}, {4, 16, 1.0f, PI32 / 12.0f}}; /* modulate.rs:88-91 */
        return modem_const_apsk(A, 4, rings, 2, out);
    }
    return MODEM_ERR_UNSUPPORTED;
}

const float* modem_lowpass_taps(size_t* n)
{
    /* src/bin/demodulate.rs:82-147 (each literal rounds to the nearest binary32, as in Rust) */
    static const float full[64] = {
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
    if (n) *n = 64;
    return full;
}

const float* modem_hilbert_taps(size_t* n)
{
    /* src/bin/demodulate.rs:48-72 (Hilbert transformer "generated with matlab") */
    static const float h[23] = {-0.007576f, -2.803e-16f, -0.019824f, 3.7096e-16f, -0.044089f, 1.3201e-16f,
                                -0.089244f, -3.2694e-16f, -0.18728f, -1.6739e-16f, -0.62794f, 0.0f,
                                0.62794f, 1.6739e-16f, 0.18728f, 3.2694e-16f, 0.089244f, -1.3201e-16f,
                                0.044089f, -3.7096e-16f, 0.019824f, 2.803e-16f, 0.007576f};
    if (n) *n = 23;
    return h;
}

int modem_phasor_by_name(const char* name, size_t br, size_t sr, modem_phasor_t* out, uint32_t* evenodd)
{
    /* the stateful rows of src/bin/modulate.rs:74-95, AMPLITUDE = 1.0 (modulate.rs:14) */
    if (!name || !out || br == 0 || sr == 0) return MODEM_ERR_INVALID;
    modem_phasor_t p;
    memset(&p, 0, sizeof p);
    p.struct_size = sizeof p;
    p.amplitude = 1.0f;
    if (evenodd) *evenodd = 0;
    if (!strcmp(name, "bfsk")) { /* BFSK::new(Freq::new(200, sr), A) */
        p.kind = MODEM_PHASOR_BFSK;
        p.bits_per_symbol = 1;
        p.deviation = modem_sample_freq(200, sr);
    } else if (!strcmp(name, "mfsk")) { /* MFSK::new(4, Freq::new(50, sr), A, IncreaseMap) */
        p.kind = MODEM_PHASOR_MFSK;
        p.bits_per_symbol = 4;
        p.deviation = modem_sample_freq(50, sr);
        p.mfsk_increase_map = 1;
    } else if (!strcmp(name, "16cpfsk")) { /* CPFSK::new(4, rates, A, 1): Freq::new(deviation * baud / 2, sr) (cpfsk.rs:19-20) */
        p.kind = MODEM_PHASOR_CPFSK;
        p.bits_per_symbol = 4;
        p.deviation = modem_sample_freq(1 * br / 2, sr);
    } else if (!strcmp(name, "msk")) { /* MSK::new(A, sps), fed by EvenOddOffset (modulate.rs:101-105) */
        if (modem_samples_per_symbol(br, sr) % 2) return MODEM_ERR_INVALID; /* msk.rs:13 assert */
        p.kind = MODEM_PHASOR_MSK;
        p.bits_per_symbol = 2;
        if (evenodd) *evenodd = 1;
    } else if (!strcmp(name, "dqpsk")) { /* DMPSK::new(2, A, PI/4, PI/2) */
        p.kind = MODEM_PHASOR_DMPSK;
        p.bits_per_symbol = 2;
        p.phase = PI32 / 4.0f;
        p.shift = PI32 / 2.0f;
    } else if (!strcmp(name, "dbpsk")) { /* DMPSK::new(1, A, PI/4, PI) */
        p.kind = MODEM_PHASOR_DMPSK;
        p.bits_per_symbol = 1;
        p.phase = PI32 / 4.0f;
        p.shift = PI32;
    } else {
        return MODEM_ERR_UNSUPPORTED;
    }
    *out = p;
    return (int)p.bits_per_symbol;
}

int modem_rrc_taps(float* out, size_t span, size_t sps, double beta)
{
    if (!out || span == 0 || sps == 0 || beta < 0.0 || beta > 1.0) return MODEM_ERR_INVALID;
    const size_t n = span * sps + 1;
    const double pi = 3.14159265358979323846;
    double energy = 0.0;
    /* two passes so no heap is needed: first energy, then scaled write-out */
    for (int pass = 0; pass < 2; ++pass) {
        const double g = pass ? 1.0 / sqrt(energy) : 1.0;
        for (size_t k = 0; k < n; ++k) {
            double t = ((double)k - (double)(n - 1) / 2.0) / (double)sps, v;
            if (fabs(t) < 1e-12) {
                v = 1.0 - beta + 4.0 * beta / pi;
            } else if (beta > 0.0 && fabs(fabs(t) - 1.0 / (4.0 * beta)) < 1e-9) {
                v = (beta / sqrt(2.0)) *
                    ((1.0 + 2.0 / pi) * sin(pi / (4.0 * beta)) + (1.0 - 2.0 / pi) * cos(pi / (4.0 * beta)));
            } else {
                double a = 4.0 * beta * t;
                v = (sin(pi * t * (1.0 - beta)) + a * cos(pi * t * (1.0 + beta))) / (pi * t * (1.0 - a * a));
            }
            if (pass) out[k] = (float)(v * g);
            else energy += v * v;
        }
    }
    return MODEM_OK;
}

float modem_sigma_for_ebn0(const modem_cfg_t* cfg, double ebn0_db)
{
    if (!cfg || !cfg->const_iq || !cfg->rx_taps || cfg->bits_per_symbol == 0) return -1.0f;
    const size_t np = (size_t)1 << cfg->bits_per_symbol;
    double es = 0.0, eh = 0.0;
    for (size_t i = 0; i < np; ++i)
        es += (double)cfg->const_iq[2 * i] * cfg->const_iq[2 * i] + (double)cfg->const_iq[2 * i + 1] * cfg->const_iq[2 * i + 1];
    es /= (double)np;
    for (uint32_t i = 0; i < cfg->n_rx_taps; ++i) eh += (double)cfg->rx_taps[i] * cfg->rx_taps[i];
    const double n0 = es / ((double)cfg->bits_per_symbol * pow(10.0, ebn0_db / 10.0));
    return (float)((double)cfg->slicer_gain * sqrt(n0 / eh) / (double)cfg->rx_gain);
}

} /* extern "C" */
