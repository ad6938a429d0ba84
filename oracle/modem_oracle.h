/*
 * modem_oracle.h -- CPU ORACLE for the modulate -> (AWGN) -> demodulate sample path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * (ramtej/rust-modem) scalar, per-sample, iterator-style algorithm, kept in the
 * reference's own structure so that it doubles as the timed CPU baseline.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.  The product (rust-modem_b200/) never links, imports or calls it.
 *
 * PARITY PINNING.  The reference is Rust (2016 nightly) and there is no Rust
 * toolchain in this image, so the reference itself cannot be run here (no
 * oracle/_ref).  The oracle is pinned against every known-answer vector the
 * reference's own unit tests hold for this path (tests/test_oracle_kat.py):
 *   data.rs:195-209 test_symbol_clock, :212-224 test_bits, :227-246 test_evenodd,
 *   digital/util.rs:22-33 test_b2b/test_max_symbol, digital/mpsk.rs:50-63 test_mpsk,
 *   digital/qam.rs:69-84 test_qam, digital/dmpsk.rs:51-84 test_dmpsk.
 * Carrier, mixer, FIR and Demodulator have NO reference test: for those the oracle
 * is pinned only by being a line-by-line restatement of the cited lines
 * ("parity unpinned" for them), cross-checked against an independent numpy model.
 * Stages the reference does not have at all (TX pulse-shaping FIR, symbol-timing
 * decimation, slicer, Philox AWGN, BER count) are DEFINED here (see "extensions").
 *
 * All citations are relative to /root/reference/.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off, glibc libm sinf/cosf, which is
 * what Rust's f32::sin/cos lower to).
 */
#ifndef MODEM_ORACLE_H
#define MODEM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- src/modem/util.rs, freq.rs, rates.rs, carrier.rs ------------------------ */
float orc_mod_trig(float x);                          /* util.rs:3-6 */
float orc_ang_freq(size_t hz);                        /* freq.rs:19-21 */
float orc_sample_freq(size_t hz, size_t sr);          /* freq.rs:24-26 */
size_t orc_samples_per_symbol(size_t br, size_t sr);  /* rates.rs:12-18 */

typedef struct {
    float sample_freq;
    size_t sample;
} orc_carrier_t;                                      /* carrier.rs:4-7 */
void orc_carrier_new(orc_carrier_t* c, size_t hz, size_t sr); /* carrier.rs:10-15 */
float orc_carrier_next(orc_carrier_t* c);                     /* carrier.rs:17-26 */

/* ---- src/modem/data.rs -------------------------------------------------------- */
typedef struct {
    size_t samples_per_symbol;
    size_t counter;
} orc_symbol_clock_t;                                 /* data.rs:14-17 */
void orc_symbol_clock_new(orc_symbol_clock_t* c, size_t sps); /* data.rs:20-25 */
int orc_symbol_clock_next(orc_symbol_clock_t* c);             /* data.rs:27-32 */

enum { ORC_CHANGED = 0, ORC_UNCHANGED = 1, ORC_FINISHED = 2 }; /* data.rs:4-8 */
typedef struct {
    int kind;
    const uint8_t* bits; /* slice of bits_per_symbol bytes, valid unless FINISHED */
    size_t len;
} orc_update_t;

typedef struct {
    const uint8_t* bits;
    size_t nbits;
    orc_symbol_clock_t clock;
    size_t bits_per_symbol;
    size_t idx;
    /* EvenOddOffset wrapper (data.rs:81-123); enabled when evenodd != 0 */
    int evenodd;
    orc_symbol_clock_t half_clock;
    uint8_t cur[2];
} orc_source_t;
void orc_bits_new(orc_source_t* s, const uint8_t* bits, size_t nbits, size_t sps, size_t bps); /* data.rs:43-52 */
void orc_evenodd_new(orc_source_t* s, const uint8_t* bits, size_t nbits, size_t sps, size_t bps); /* data.rs:88-99 */
orc_update_t orc_source_next(orc_source_t* s);        /* data.rs:66-79, 102-122 */

/* ---- src/modem/digital/ ------------------------------------------------------- */
float orc_bit_to_sign(uint8_t b);                     /* digital/util.rs:1-3 */
uint8_t orc_bytes_to_bits(const uint8_t* b, size_t n);/* digital/util.rs:5-11 */
size_t orc_max_symbol(size_t bits_per_symbol);        /* digital/util.rs:13-15 */

enum {
    ORC_BASK = 0, ORC_BPSK, ORC_QPSK, ORC_QAM, ORC_MPSK, ORC_OQPSK, ORC_DCQPSK, ORC_APSK,
    ORC_BFSK, ORC_MFSK_DEFAULT, ORC_MFSK_INCREASE, ORC_CPFSK, ORC_MSK, ORC_DMPSK
};

typedef struct {
    uint8_t start, end;   /* Range<u8> */
    float radius, phase;
} orc_ring_t;                                         /* digital/apsk.rs:60-67 */

typedef struct {
    int scheme;
    size_t bits_per_symbol;
    float amplitude;      /* as stored by each ::new (already scaled where the reference scales it) */
    float phase;          /* BPSK phase / MPSK phase_offset / DMPSK phase / BFSK phase / MFSK phase_offset */
    float phase_cos, phase_sin; /* QPSK, QAM */
    float max_symbol;     /* QAM */
    size_t bits_per_carrier; /* QAM */
    float num_symbols;    /* MPSK */
    int even;             /* DCQPSK */
    orc_ring_t rings[8];  /* APSK */
    size_t n_rings;
    float deviation;      /* BFSK, MFSK: deviation.sample_freq(); CPFSK: freq */
    uint8_t prev;         /* BFSK */
    float cur_coef;       /* MFSK */
    int max_symbol_i;     /* MFSK DefaultMap */
    float shift;          /* DMPSK */
    size_t samples_per_bit; /* MSK */
} orc_phasor_t;

void orc_bask_new(orc_phasor_t* p, float a);                                   /* bask.rs:8-12 */
void orc_bpsk_new(orc_phasor_t* p, float phase, float amplitude);              /* bpsk.rs:10-15 */
void orc_qpsk_new(orc_phasor_t* p, float phase, float amplitude);              /* qpsk.rs:11-17 */
void orc_qam_new(orc_phasor_t* p, size_t bps, float phase, float amplitude);   /* qam.rs:15-30 */
void orc_mpsk_new(orc_phasor_t* p, size_t bps, float phase_offset, float amplitude); /* mpsk.rs:14-21 */
void orc_oqpsk_new(orc_phasor_t* p, float amplitude);                          /* oqpsk.rs:9-13 */
void orc_dcqpsk_new(orc_phasor_t* p, float amplitude);                         /* dcqpsk.rs:16-21 */
int  orc_apsk_new(orc_phasor_t* p, float amplitude, size_t bps, const orc_ring_t* rings, size_t n); /* apsk.rs:25-33; 0 if verify() fails */
void orc_bfsk_new(orc_phasor_t* p, size_t dev_hz, size_t sr, float a);         /* bfsk.rs:14-21 */
void orc_mfsk_new(orc_phasor_t* p, size_t bps, size_t dev_hz, size_t sr, float a, int increase_map); /* mfsk.rs:47-58 */
void orc_cpfsk_new(orc_phasor_t* p, size_t bps, size_t br, size_t sr, float a, size_t deviation); /* cpfsk.rs:14-23 */
void orc_msk_new(orc_phasor_t* p, float a, size_t sps);                        /* msk.rs:12-19 */
void orc_dmpsk_new(orc_phasor_t* p, size_t bps, float a, float phase, float shift); /* dmpsk.rs:16-23 */
/* The 13 memoryless/stateful `-m` names of src/bin/modulate.rs:74-95 with its constants. */
int  orc_phasor_by_name(orc_phasor_t* p, const char* name, size_t br, size_t sr);

void  orc_phasor_update(orc_phasor_t* p, size_t s, const uint8_t* b);          /* digital/phasor.rs:4 + overrides */
float orc_phasor_i(const orc_phasor_t* p, size_t s, const uint8_t* b);         /* digital/phasor.rs:6 */
float orc_phasor_q(const orc_phasor_t* p, size_t s, const uint8_t* b);         /* digital/phasor.rs:7 */

/* ---- src/modem/fir.rs --------------------------------------------------------- */
typedef struct {
    const float* coefs;
    size_t n;
    float* history;
    size_t idx;
} orc_fir_t;                                          /* fir.rs:3-7 */
int orc_fir_new(orc_fir_t* f, const float* coefs, size_t n); /* fir.rs:10-16 (allocates history) */
float orc_fir_add(orc_fir_t* f, float sample);               /* fir.rs:18-34 */
void orc_fir_free(orc_fir_t* f);

/* ---- src/modem/modulator.rs ---------------------------------------------------- */
typedef struct {
    float carrier, i, q;
} orc_iq_sample_t;                                    /* modulator.rs:22-26 */
void orc_iq_modulate(const orc_iq_sample_t* s, float* re, float* im); /* modulator.rs:37-48 */
/* DigitalModulator::next (modulator.rs:85-100); returns 0 on Finished */
int orc_digital_modulator_next(orc_carrier_t* c, orc_phasor_t* p, orc_source_t* src, orc_iq_sample_t* out);

/* ---- src/modem/pll.rs, demodulator.rs ------------------------------------------ */
typedef struct {
    float phase_offset;
} orc_pll_t;                                          /* pll.rs:5-7 */
void orc_pll_handle(orc_pll_t* pll, float carrier_phase, float x_re, float x_im); /* pll.rs:16-22 */

typedef struct {
    orc_carrier_t carrier;
    orc_pll_t pll;
    orc_fir_t lpi, lpq;
} orc_demod_t;                                        /* demodulator.rs:7-15 */
int orc_demod_new(orc_demod_t* d, orc_carrier_t carrier, const float* taps, size_t n); /* demodulator.rs:20-30 */
void orc_demod_lock_step(orc_demod_t* d, float x_re, float x_im);   /* one iteration of lock_phase, demodulator.rs:32-36 */
void orc_demod_next(orc_demod_t* d, float x_re, float* i, float* q);/* demodulator.rs:44-55 */
void orc_demod_free(orc_demod_t* d);

/* ---- src/bin/demodulate.rs tap tables ------------------------------------------ */
const float* orc_lowpass_taps(size_t* n);             /* demodulate.rs:82-147 (64 taps) */
const float* orc_hilbert_taps(size_t* n);             /* demodulate.rs:48-72 (23 taps) */

/* =============================== extensions ======================================
 * Not in the reference; DEFINED here (SURVEY.md section 8c "extension semantics").
 */

/* Root-raised-cosine taps: n = span*sps + 1, roll-off beta, computed in binary64,
 * normalised to unit energy, rounded to binary32. */
void orc_rrc_taps(float* out, size_t span, size_t sps, double beta);

/* Philox4x32-10 (Salmon et al., SC'11).  ctr/key/out are 4/2/4 words. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- AWGN (extension; nothing in the reference).  DEFINED here as an explicit sequence of IEEE binary32 operations --
 * mul, add, fused multiply-add (fmaf), sqrtf, all correctly rounded, evaluated in the order written -- so that a CPU and
 * a GPU produce the same bits without either having to imitate the other's math library (round 1 defined the noise through
 * glibc's logf/sinf/cosf and paid a binary64 libm per sample on the device to stay bit-identical).
 *
 *   normal pair from two Philox words (r0, r1)                                     [orc_box_muller]
 *     u  = ((float)(r0 >> 9) + 0.5f) * 2^-23           uniform on (0, 1), exact in binary32
 *     ln u:  ix = bits(u) - 0x3f3504f3;  e = (int)ix >> 23;  m = float(bits = (ix & 0x7fffff) + 0x3f3504f3)
 *            f = m - 1.0f                                in [sqrt(1/2) - 1, sqrt(2) - 1), exact
 *            q = L6; q = fmaf(q, f, L5); ... q = fmaf(q, f, L0)
 *            lnm = f * fmaf(f, q, 1.0f);   lnu = fmaf((float)e, LN2, lnm)
 *     rad = sqrtf(-2.0f * lnu)
 *     angle: j = r1 >> 8 (24 bits = fraction of a turn); oct = j >> 21; k = j & 0x1fffff; if (oct & 1) k = 0x200000 - k
 *            x = (float)k * ANG                          in [0, pi/4]
 *            z = x * x;  s = fmaf(x * z, fmaf(fmaf(S2, z, S1), z, S0), x);  c = fmaf(z, fmaf(fmaf(fmaf(C3, z, C2), z, C1), z, C0), 1.0f)
 *            octants 1, 2, 5, 6 swap (s, c); the cosine is negated in octants 2..5, the sine in octants 4..7
 *     n0 = rad * cos, n1 = rad * sin
 *   constants (binary32, hexadecimal): L0..L6 = -0x1.00001cp-1, 0x1.555802p-2, -0x1.ffa938p-3, 0x1.97ecccp-3, -0x1.5e404cp-3,
 *     0x1.495358p-3, -0x1.ab64d2p-4;  LN2 = 0x1.62e43p-1;  ANG = 0x1.921fb6p-22 (2 pi 2^-24);  S0..S2 = -0x1.555552p-3,
 *     0x1.110c2ap-7, -0x1.9aca02p-13;  C0..C3 = -0x1p-1, 0x1.55554cp-5, -0x1.6c0e0cp-10, 0x1.9a6fd8p-16.
 *     (|ln error| < 5e-7 relative, |sin, cos error| < 6e-8: the distribution is the algorithm's by definition; its
 *     closeness to a Gaussian is what tests/test_oracle_kat.py and the BER-against-theory tests check.)
 *
 *   noise of sample n of global frame g, seed s                                    [orc_awgn_sample]
 *     one Philox4x32-10 block per aligned QUAD of samples and per rail: key = (s lo, s hi), counter =
 *     (n/4 lo, (n/4 hi) | rail << 31, g lo, g hi), rail 0 = real parts, rail 1 = imaginary parts.  Words (0,1) ->
 *     normal pair for samples 4q, 4q+1; words (2,3) -> pair for samples 4q+2, 4q+3.  x.re += sigma * a, x.im += sigma * b
 *     (a product and a sum, each rounded).  The demodulator reads real parts only (demodulator.rs:45-48), so a receiver
 *     that adds the noise on the fly needs ONE block per four samples and wastes none of it.
 */
void orc_box_muller(uint32_t r0, uint32_t r1, float* n0, float* n1);
void orc_awgn_sample(uint64_t seed, uint64_t frame, uint64_t n, float sigma, float* re, float* im);
void orc_awgn(float* buf /*[F][L][2]*/, size_t F, size_t L, float sigma, uint64_t seed, uint64_t frame0);

/* ---- random payload bits (extension): bit j of global frame g = bit (j % 32) of word (j % 128) / 32 of the Philox4x32-10
 * block with key = (s lo, s hi ^ 0x62697473) and counter = (j/128 lo, j/128 hi, g lo, g hi); one byte (0/1) per bit. */
void orc_random_bits(uint8_t* bits /*[F][nbits]*/, size_t F, size_t nbits, uint64_t seed, uint64_t frame0);

/* ---- packed payload (extension, modem_gpu_loopback_packed): rows of ceil(nbits/8) bytes; bit j of a frame = bit 7 - j%8 of
 * byte j/8 (first bit most significant: the order in which bytes_to_bits, digital/util.rs:5-11, packs a symbol); pad bits
 * of a row's last byte are written as zero and ignored on input.  unpack yields the one-byte-per-bit rows of
 * data.rs:35-40, pack is its inverse on bit 0 of every byte. */
void orc_unpack_bits(const uint8_t* packed /*[F][ceil(nbits/8)]*/, uint8_t* bits /*[F][nbits]*/, size_t F, size_t nbits);
void orc_pack_bits(const uint8_t* bits /*[F][nbits]*/, uint8_t* packed /*[F][ceil(nbits/8)]*/, size_t F, size_t nbits);

/* Es = mean |c|^2 of the scheme's constellation; sigma for a given Eb/N0 so that the
 * slicer sees the textbook SNR (DESIGN.md "AWGN scaling"). */
float orc_sigma_for_ebn0(const float* const_iq, size_t n_points, size_t bps, float slicer_gain,
                         float rx_gain, const float* rx_taps, size_t n_rx, double ebn0_db);

/* Whole-path batch drivers: each frame is run through the streaming objects above,
 * exactly as a `src/bin`-style caller would compose them. */
typedef struct {
    /* mapper */
    char scheme[16];          /* a modulate.rs -m name, e.g. "qpsk" */
    /* rates / carrier: Rates::new(br, sr), Carrier::new(Freq::new(cf, sr)) */
    size_t baud_rate, sample_rate, carrier_hz;
    size_t sample0;           /* Carrier.sample when the frame starts (0 = fresh Carrier::new) */
    /* TX shaping extension: n_tx_taps == 0 => rectangular hold (exact reference) */
    const float* tx_taps;
    size_t n_tx_taps;
    /* RX */
    const float* rx_taps;
    size_t n_rx_taps;
    float phase_offset;       /* PLL::phase_offset held constant (0 in coherent loopback) */
    /* decision extension */
    size_t decision_delay;
    float slicer_gain;
} orc_path_t;

size_t orc_frame_samples(const orc_path_t* p, size_t nbits);      /* floor(nbits/bps)*sps */
size_t orc_decided_symbols(const orc_path_t* p, size_t L);        /* 0 if L <= delay else (L-1-delay)/sps+1 */
size_t orc_bits_per_symbol(const orc_path_t* p);
/* constellation the slicer uses: n_tables * 2^bps (i,q) pairs, evaluated through the
 * phasor's own i()/q() (dcqpsk: table 0 = first symbol's constellation). Returns n_tables. */
size_t orc_constellation(const orc_path_t* p, float* out_iq, size_t cap_points);

/* bits [F][nbits] (one byte per bit) -> tx [F][L] complex (interleaved re,im), optional
 * iq [F][L] baseband (i,q) before mixing.  Returns 0 on success. */
int orc_modulate(const orc_path_t* p, const uint8_t* bits, size_t F, size_t nbits,
                 float* tx, float* iq);
/* rx [F][L] complex -> filt [F][L] (I,Q) full rate (nullable), sym [F][K] (nullable),
 * bits_out [F][K*bps] (nullable). */
int orc_demodulate(const orc_path_t* p, const float* rx, size_t F, size_t L,
                   float* filt, uint8_t* sym, uint8_t* bits_out);
/* fused driver used as the CPU baseline: modulate -> (AWGN if sigma>0) -> demodulate ->
 * count bit errors over decided symbols; frames sharded over `threads` pthreads.
 * counters[0] += bit errors, counters[1] += bits compared. */
int orc_loopback(const orc_path_t* p, const uint8_t* bits, size_t F, size_t nbits,
                 float sigma, uint64_t seed, uint64_t frame0, int threads,
                 uint8_t* sym /*nullable [F][K]*/, uint8_t* bits_out /*nullable*/,
                 uint64_t counters[2]);


/* ---- src/bin drivers: the two binaries' sample paths, composed from the objects above ---- */
/* src/bin/modulate.rs:118-133 without --iq: `preamble` samples of Modulator + phasor::Raw(amplitude)
 * (modulator.rs:51-62, phasor.rs:5-24) mapped through `x.modulate().re`, then the DigitalModulator's
 * samples `.re`, all on ONE Carrier (the counter carries over, modulate.rs:120,128).
 * out [F][preamble + L] f32. */
int orc_modulate_real(const orc_path_t* p, const uint8_t* bits, size_t F, size_t nbits, size_t preamble,
                      float preamble_amplitude, float* out);
/* Complex samples of the same tone: tx [F][n][2]. */
int orc_preamble(const orc_path_t* p, size_t F, size_t n, float amplitude, float* tx);
/* src/bin/demodulate.rs:29-43 per frame: x = sample as f32 (real input x[F][L]); analytic = (x, hilbert.add(x));
 * Demodulator::new(carrier, analytic, lowpass); lock_phase() over the first `lock` samples
 * (demodulator.rs:32-36); then Demodulator::next for the remaining L - lock samples.
 * analytic_im (nullable) [F][L]: caller-supplied imaginary parts instead of the Hilbert FIR.
 * po_out [F] (nullable) = PLL.phase_offset after the lock; filt [F][L-lock][2] (nullable) the (I,Q) stream the
 * binary prints; sym/bits_out (nullable) the decimator/slicer extension over that stream.
 * lock == 0: no lock, phase offset = p->phase_offset.  Returns -3 if L < lock (the reference's unwrap panic). */
int orc_demodulate_real(const orc_path_t* p, const float* x, const float* analytic_im, size_t F, size_t L, size_t lock,
                        const float* hilbert, size_t n_hilbert, float* po_out, float* filt, uint8_t* sym, uint8_t* bits_out);

#ifdef __cplusplus
}
#endif
#endif
