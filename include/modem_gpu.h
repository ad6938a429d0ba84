/*
 * modem_gpu.h -- C ABI of the B200 (sm_100a) batched modulate -> (AWGN) -> demodulate path.
 *
 * The reference (ramtej/rust-modem, paths below relative to /root/reference/) has no FFI
 * or plugin interface; the boundary it offers is the public Rust API of crate `modem`
 * as used by src/bin/{modulate,demodulate}.rs.  This header is what a thin Rust module
 * (`modem::gpu`, see INTEGRATION.md and rust-modem_b200/rust/) binds with `extern "C"`,
 * what the C++ mirror of that API (rust-modem_b200/host/modem.hpp) calls, and what the
 * Python tests/bench call through ctypes -- the same symbols for all three.
 *
 * Conventions
 *  - plain pointers and sizes only; every entry returns int (0 = MODEM_OK, <0 = error);
 *    nothing throws or aborts across this boundary.  The Rust side turns a non-zero
 *    return into the panic!/assert! the reference would raise (e.g. modulate.rs:94).
 *  - the caller owns every buffer.  Data pointers may be DEVICE pointers (used in place)
 *    or HOST pointers (pageable or pinned; staged through the context's own device
 *    scratch with async copies on the context's streams).
 *  - one context = one device + one stream; a context is not thread-safe, separate
 *    contexts are independent.
 *  - frames are independent: the NCO sample counter restarts at cfg.sample0 for each
 *    frame (a fresh Carrier::new has sample = 0, carrier.rs:10-15) and the FIR history
 *    starts at zero (fir.rs:13).
 *  - there is no CPU fallback: every compute entry fails with MODEM_ERR_NO_DEVICE /
 *    MODEM_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MODEM_GPU_H
#define MODEM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODEM_GPU_ABI_VERSION 4

enum {
    MODEM_OK = 0,
    MODEM_ERR_INVALID = -1,     /* bad argument / configuration (reference: assert!/panic!) */
    MODEM_ERR_CUDA = -2,        /* a CUDA runtime call failed; see modem_gpu_last_error */
    MODEM_ERR_NOMEM = -3,
    MODEM_ERR_UNSUPPORTED = -4,
    MODEM_ERR_NCCL = -5,
    MODEM_ERR_NO_DEVICE = -6
};

/* num::Complex32 / float2: the sample type of modulator.rs:45-48 and demodulator.rs:8 */
typedef struct {
    float re, im;
} modem_c32_t;

typedef struct modem_ctx modem_ctx_t;
typedef struct modem_comm modem_comm_t;

/* cfg.flags */
#define MODEM_FLAG_FUSED_MAC 0x1u /* FIR accumulates with FMA instead of the reference's
                                     separate mul+add (fir.rs:23).  Faster on long taps, no
                                     longer bit-identical to the CPU path (still << 1e-5). */

#define MODEM_FLAG_NO_TMEM 0x2u   /* The tuned sps-8 RX / fused loopback kernels park frame-invariant NCO values in
                                     TENSOR MEMORY (tcgen05.alloc: 64 columns per CTA, 256 of an SM's 512 columns for the
                                     fused kernel, all 512 for the plain 64-tap RX kernel at 8 CTAs per SM).
                                     tcgen05.alloc blocks until columns are free, so next to another tensor-memory user
                                     on the same device (a tcgen05 GEMM on another stream) these kernels can serialise
                                     with it.  This flag selects instantiations that allocate no tensor memory (the
                                     values are read through L1 instead: ~10 % slower, bit-identical results). */

/*
 * Path configuration.  Field names follow the reference's parameter vocabulary.
 */
typedef struct {
    uint32_t struct_size;        /* = sizeof(modem_cfg_t); ABI guard */
    uint32_t bits_per_symbol;    /* DigitalPhasor::bits_per_symbol() (digital/phasor.rs:2), 1..8 */
    uint32_t samples_per_symbol; /* Rates::samples_per_symbol (rates.rs:16) */
    uint32_t n_tables;           /* 1; 2 for dcqpsk (dcqpsk.rs:42-44): symbol k uses table k % n_tables */
    const float* const_iq;       /* [n_tables][2^bps][2] (i,q) per symbol index (MSB-first bits,
                                    digital/util.rs:5-11), from the mapper formulas digital/<scheme>.rs */
    uint32_t q_offset;           /* EvenOddOffset (data.rs:81-123): Q bit of a symbol is applied this
                                    many samples late (sps/2 for oqpsk); 0 otherwise.  Needs bps == 2. */
    float sample_freq;           /* Freq::sample_freq() (freq.rs:24-26), radians per sample */
    float phase_offset;          /* PLL::phase_offset (pll.rs:6) held for the whole call; 0 = coherent */
    uint64_t sample0;            /* Carrier.sample at the first sample of every frame (carrier.rs:6) */
    uint32_t n_tx_taps;          /* 0 => rectangular hold, the reference's TX (data.rs:66-79) */
    const float* tx_taps;        /* else: FIR (fir.rs semantics) over the zero-stuffed symbol train */
    uint32_t n_rx_taps;          /* the `lp` filter of Demodulator::new (demodulator.rs:20-30) */
    const float* rx_taps;        /* e.g. modem_lowpass_taps() == src/bin/demodulate.rs:82-147 */
    uint32_t decision_delay;     /* symbol k is sliced at sample k*sps + decision_delay */
    float rx_gain;               /* 2.0 (demodulator.rs:53-54) */
    float slicer_gain;           /* end-to-end gain g: slicer compares against g * const_iq */
    uint32_t flags;              /* MODEM_FLAG_* */
} modem_cfg_t;

/* ------------------------------------------------------------------ host-side helpers
 * Pure host code (no device needed): the constants a caller of the reference would get
 * from Freq / Rates / the digital::* constructors, computed the reference's way (binary32,
 * unfused, glibc sinf/cosf).
 */
float modem_sample_freq(size_t hz, size_t sr);                 /* freq.rs:19-26 */
size_t modem_samples_per_symbol(size_t baud_rate, size_t sample_rate); /* rates.rs:12-18 */

/* Constellation tables; each writes 2^bps (i,q) pairs to out_iq and returns bps (<0 on error). */
int modem_const_bask(float amplitude, float* out_iq);                                  /* bask.rs */
int modem_const_bpsk(float phase, float amplitude, float* out_iq);                     /* bpsk.rs */
int modem_const_qpsk(float phase, float amplitude, float* out_iq);                     /* qpsk.rs */
int modem_const_qam(uint32_t bps, float phase, float amplitude, float* out_iq);        /* qam.rs */
int modem_const_mpsk(uint32_t bps, float phase_offset, float amplitude, float* out_iq);/* mpsk.rs */
int modem_const_oqpsk(float amplitude, float* out_iq);                                 /* oqpsk.rs */
int modem_const_dcqpsk(float amplitude, float* out_iq /* 2 tables: 8 pairs */);        /* dcqpsk.rs */
typedef struct {
    uint8_t start, end; /* Range<u8> */
    float radius, phase;
} modem_ring_t;                                                                        /* apsk.rs:60-67 */
int modem_const_apsk(float amplitude, uint32_t bps, const modem_ring_t* rings, size_t n_rings,
                     float* out_iq);                                                   /* apsk.rs */
/* The memoryless `-m` names of src/bin/modulate.rs:74-95 with that file's constants
 * (AMPLITUDE 1.0, bpsk phase pi/4, 16apsk rings, ...).  out_iq needs room for 512 pairs.
 * Returns bps; *n_tables and *evenodd (1 => use q_offset = sps/2, modulate.rs:101-107) are set.
 * Unknown or stateful (bfsk/mfsk/msk/cpfsk/dmpsk) names return MODEM_ERR_UNSUPPORTED. */
int modem_const_by_name(const char* name, float* out_iq, uint32_t* n_tables, uint32_t* evenodd);

const float* modem_lowpass_taps(size_t* n);  /* src/bin/demodulate.rs:82-147, 64 taps */
const float* modem_hilbert_taps(size_t* n);  /* src/bin/demodulate.rs:48-72, 23 taps */

/* Stateful / time-varying phasors (digital/{bfsk,mfsk,cpfsk,msk,dmpsk}.rs): (i,q) depends on the sample
 * counter and, for bfsk/mfsk/dmpsk, on a phase carried from symbol to symbol (DigitalPhasor::update).
 * The phasor sees s = Carrier.sample AFTER the increment of Carrier::next (modulator.rs:86-97), i.e.
 * s = cfg.sample0 + n + 1 for sample n of a frame; every frame starts from the constructor's state. */
enum {
    MODEM_PHASOR_TABLE = 0, /* the memoryless constellation table of modem_cfg_t.const_iq */
    MODEM_PHASOR_BFSK = 1,  /* bfsk.rs:23-56  */
    MODEM_PHASOR_MFSK = 2,  /* mfsk.rs:60-84  */
    MODEM_PHASOR_CPFSK = 3, /* cpfsk.rs:25-44 */
    MODEM_PHASOR_MSK = 4,   /* msk.rs:21-36 (fed by EvenOddOffset: cfg.q_offset = sps/2, modulate.rs:101-107) */
    MODEM_PHASOR_DMPSK = 5  /* dmpsk.rs:29-42 */
};
typedef struct {
    uint32_t struct_size;     /* = sizeof(modem_phasor_t) */
    uint32_t kind;            /* MODEM_PHASOR_* */
    uint32_t bits_per_symbol; /* must equal cfg.bits_per_symbol */
    float amplitude;
    float deviation;          /* bfsk.rs:16, mfsk.rs:52: deviation.sample_freq(); cpfsk.rs:19-20: freq */
    float phase;              /* dmpsk.rs:20 initial phase */
    float shift;              /* dmpsk.rs:21 */
    uint32_t mfsk_increase_map; /* mfsk.rs: 0 = DefaultMap (:22-27), 1 = IncreaseMap (:31-35) */
} modem_phasor_t;
/* The stateful `-m` names of src/bin/modulate.rs:74-95 (bfsk, mfsk, msk, 16cpfsk, dqpsk, dbpsk) with that
 * file's constants.  Returns bps; *evenodd = 1 for msk.  Other names: MODEM_ERR_UNSUPPORTED. */
int modem_phasor_by_name(const char* name, size_t baud_rate, size_t sample_rate, modem_phasor_t* out, uint32_t* evenodd);
/* root-raised-cosine, span*sps+1 taps, unit energy (extension) */
int modem_rrc_taps(float* out, size_t span, size_t sps, double beta);
/* AWGN sigma that gives the slicer the textbook Eb/N0 (DESIGN.md "AWGN scaling") */
float modem_sigma_for_ebn0(const modem_cfg_t* cfg, double ebn0_db);

/* ------------------------------------------------------------------ context */
int modem_gpu_device_count(int* n);
int modem_gpu_create(modem_ctx_t** ctx, int device, const modem_cfg_t* cfg);
void modem_gpu_destroy(modem_ctx_t* ctx);
/* Use the caller's CUDA stream (a cudaStream_t / CUstream handle) for all work of this ctx. */
int modem_gpu_set_stream(modem_ctx_t* ctx, void* cuda_stream);
/* Multi-carrier bank: frame f uses sample_freq[f / frames_per_channel] and
 * phase_offset[...] (nullable => cfg.phase_offset).  n_channels == 0 restores cfg.sample_freq. */
int modem_gpu_set_channels(modem_ctx_t* ctx, size_t n_channels, const float* sample_freq,
                           const float* phase_offset, size_t frames_per_channel);
int modem_gpu_synchronize(modem_ctx_t* ctx);
/* Replace the table mapper of modem_gpu_modulate* by a stateful phasor (NULL or kind TABLE restores the
 * table).  TX only: the reference has no receiver for these schemes. */
int modem_gpu_set_phasor(modem_ctx_t* ctx, const modem_phasor_t* phasor);

size_t modem_gpu_frame_samples(const modem_ctx_t* ctx, size_t nbits);   /* floor(nbits/bps)*sps (data.rs:54-63) */
size_t modem_gpu_decided_symbols(const modem_ctx_t* ctx, size_t L);     /* floor((L-1-delay-q_offset)/sps)+1 */

/* ------------------------------------------------------------------ the path
 * modulate: bits [F][nbits], one byte per bit (data.rs:35-40)  ->  tx [F][L] complex,
 *   L = modem_gpu_frame_samples(nbits).  Replaces, per sample, Bits::next (data.rs:66-79),
 *   DigitalPhasor::next (digital/phasor.rs:9-11), Carrier::next (carrier.rs:21-26) and
 *   IQSample::modulate (modulator.rs:37-48) -- i.e. `DigitalModulator::new(..).map(|s| s.modulate())`.
 *   iq (nullable): the baseband (i,q) pairs before mixing (IQSample.i/.q, modulate.rs:110-113).
 */
int modem_gpu_modulate(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits,
                       modem_c32_t* tx, modem_c32_t* iq);

/* preamble: the carrier sync tone of modulate.rs:118-126 -- `Modulator::new(&mut carrier,
 *   Box::new(phasor::Raw::new(amplitude)))` (modulator.rs:8-20,51-62, phasor.rs:5-24) mapped through
 *   IQSample::modulate: tx [F][n] complex, sample counter from cfg.sample0. */
int modem_gpu_preamble(modem_ctx_t* ctx, size_t F, size_t n, float amplitude, modem_c32_t* tx);

/* modulate_real: exactly the f32 stream src/bin/modulate.rs writes without --iq (modulate.rs:118-133):
 *   `preamble` samples of the sync tone (amplitude preamble_amplitude; 0 samples = none), then the data
 *   samples, each `x.modulate().re`.  The Carrier is shared: its sample counter runs on from the tone into
 *   the data (modulate.rs:120,128), so data sample n uses Carrier.sample = cfg.sample0 + preamble + n.
 *   out [F][preamble + L] f32, L = modem_gpu_frame_samples(nbits). */
int modem_gpu_modulate_real(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits, size_t preamble,
                            float preamble_amplitude, float* out);

/* awgn (extension): buf[f][n] += sigma * N(0,1) per component.  Philox4x32-10 keyed by `seed`, one block per aligned quad
 * of samples and per rail (counter = (n/4, rail in bit 63, global frame id frame0+f)), Box-Muller as an explicit
 * sequence of correctly rounded binary32 operations: the definition is in oracle/modem_oracle.h ("AWGN") and is what a
 * CPU reproduces bit for bit.  In place. */
int modem_gpu_awgn(modem_ctx_t* ctx, modem_c32_t* buf, size_t F, size_t L, float sigma,
                   uint64_t seed, uint64_t frame0);

/* random_bits (extension, BASELINE config 4 "bits from Philox too"): payload bits for F frames generated on the device,
 * bits [F][nbits] one byte (0/1) per bit (host or device pointer); bit j of global frame frame0+f is bit j % 32 of word
 * (j % 128) / 32 of the Philox4x32-10 block with key (seed lo, seed hi ^ 0x62697473), counter (j/128, frame id). */
int modem_gpu_random_bits(modem_ctx_t* ctx, uint8_t* bits, size_t F, size_t nbits, uint64_t seed, uint64_t frame0);

/* demodulate: rx [F][L] complex -> per frame K = modem_gpu_decided_symbols(L) decisions.
 *   Replaces Demodulator::next (demodulator.rs:44-55) with its two FIRFilter::add
 *   (fir.rs:18-34), plus the decimator/slicer extension.
 *   sym  [F][K]      (nullable) symbol indices
 *   bits [F][K*bps]  (nullable) demapped bits, one byte per bit, MSB first
 *   soft [F][K]      (nullable) the filtered (I,Q) at the decision instants
 *   filt [F][L]      (nullable) the full-rate filtered (I,Q) stream the reference emits
 *                    (compute-bound parity/debug output)
 *   sigma > 0 adds the same noise as modem_gpu_awgn(seed, frame0) while loading rx
 *   (the buffer itself is not modified).
 */
int modem_gpu_demodulate(modem_ctx_t* ctx, const modem_c32_t* rx, size_t F, size_t L,
                         uint8_t* sym, uint8_t* bits, modem_c32_t* soft, modem_c32_t* filt,
                         float sigma, uint64_t seed, uint64_t frame0);

/* Sample formats of the carrier-recovery front end (src/bin/demodulate.rs:29-34) */
enum {
    MODEM_SAMPLES_C32 = 0, /* analytic signal supplied by the caller: num::Complex<f32> per sample */
    MODEM_SAMPLES_F32 = 1, /* real f32 samples (what modulate writes); analytic = (x, hilbert.add(x)) */
    MODEM_SAMPLES_I16 = 2  /* real native-endian i16 samples (src/bin/util.rs:3-37), `x as f32` */
};
/* lock_phase: Demodulator::lock_phase (demodulator.rs:32-36): for each frame, the first lock_samples
 *   (LOCK_SAMPLES = 64, demodulator.rs:5) analytic samples drive PLL::handle (pll.rs:16-22) with
 *   Carrier::next() from cfg.sample0; phase_offset [F] (host or device) receives PLL.phase_offset.
 *   For real formats the imaginary part is the Hilbert FIR (fir.rs semantics, zero history) of the
 *   real samples (demodulate.rs:31-34); hilbert_taps == NULL uses modem_hilbert_taps(). */
int modem_gpu_lock_phase(modem_ctx_t* ctx, const void* samples, uint32_t fmt, size_t F, size_t L,
                         const float* hilbert_taps, size_t n_hilbert, size_t lock_samples, float* phase_offset);

/* demodulate_real: the whole path of src/bin/demodulate.rs:29-43 per frame: lock_phase over the first
 *   lock_samples samples (0 = no lock: cfg.phase_offset is used), then Demodulator::next over the remaining
 *   Lr = L - lock_samples samples with the carrier counter running on (cfg.sample0 + lock_samples + n), the
 *   frame's own PLL.phase_offset and fresh low-pass history.  Outputs as modem_gpu_demodulate with L := Lr
 *   (filt [F][Lr] is the `i:{}\tq:{}` stream the binary prints); phase_offset [F] (nullable) receives the
 *   locked offsets.  L < lock_samples is the reference's `unwrap()` panic: MODEM_ERR_INVALID. */
int modem_gpu_demodulate_real(modem_ctx_t* ctx, const void* samples, uint32_t fmt, size_t F, size_t L,
                              size_t lock_samples, const float* hilbert_taps, size_t n_hilbert, float* phase_offset,
                              uint8_t* sym, uint8_t* bits, modem_c32_t* soft, modem_c32_t* filt);

/* demodulate_count: the stream-ordered, device-resident form used inside a loopback: like
 * modem_gpu_demodulate (sym / bits nullable) but additionally compares every decided bit with
 * ref_bits [F][ref_stride] (the bits that were modulated) and ACCUMULATES into the DEVICE
 * counters[2] = {bit errors, bits compared}.  All pointers must be device pointers; nothing is
 * synchronised, so the counters can feed modem_gpu_allreduce_counters on the same stream. */
int modem_gpu_demodulate_count(modem_ctx_t* ctx, const modem_c32_t* rx, size_t F, size_t L,
                               uint8_t* sym, uint8_t* bits, const uint8_t* ref_bits, size_t ref_stride,
                               uint64_t* counters, float sigma, uint64_t seed, uint64_t frame0);

/* ber_sweep (extension, BASELINE config 4): Monte-Carlo BER over n_points noise levels.  The frames are
 * modulated ONCE into tx (nullable => context scratch); then for every point p the RX kernel re-reads the
 * clean TX buffer, adds AWGN of sigmas[p] on the fly (Philox stream keyed by seed + p, counter = global
 * frame id frame0 + f and sample index) and ACCUMULATES {bit errors, bits compared} into the DEVICE array
 * counters[p][2].  Stream-ordered, device pointers (sigmas is a host array).  Shard frames across ranks by
 * giving each rank its own bits and frame0, then reduce all points with ONE
 * modem_gpu_allreduce_counters(comm, counters, 2 * n_points). */
int modem_gpu_ber_sweep(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits, size_t n_points,
                        const float* sigmas, uint64_t seed, uint64_t frame0, modem_c32_t* tx,
                        uint64_t* counters);

/* loopback_device: the whole loopback, stream-ordered and device-resident (nothing is synchronised):
 * bits / tx (nullable) / sym / bits_out (nullable) are device pointers, counters is a DEVICE u64[2] that is
 * accumulated into.  The headline shape (QPSK table, rectangular hold, 8 samples per symbol, the 64-tap low-pass
 * of src/bin/demodulate.rs:82-147, odd decision delay, exact MACs, sigma == 0, no phase offset, rows of bits on
 * 8-byte boundaries) runs as ONE fused kernel that makes the TX samples, stores them to tx (not at all when tx is
 * null) and demodulates them from registers; every buffer is bit-identical to the two-kernel path.  The same holds for
 * any other samples-per-symbol count up to 64 with that table / filter (the reference's default rates, sps 45:
 * rates.rs:16 with modulate.rs:44-58; decision delay <= 63, even frame length, rows of bits on 2-byte boundaries).  All
 * other shapes take the two kernels (tx null => context scratch).  MODEM_GPU_NO_FUSED_LOOP=1 forces two kernels; MODEM_GPU_LOOP_CHUNK=n
 * cuts the two-kernel path into chunks of n frames with TX(c+1) || RX(c) on two streams (measured slower). */
int modem_gpu_loopback_device(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits,
                              float sigma, uint64_t seed, uint64_t frame0, modem_c32_t* tx,
                              uint8_t* sym, uint8_t* bits_out, uint64_t* counters);

/* loopback: modulate -> (AWGN) -> demodulate -> count bit errors against the input bits
 * over the decided symbols.  tx (nullable => context scratch) receives the clean TX
 * samples.  counters[0] += bit errors, counters[1] += bits compared (host pointer,
 * read back after the work completes; this call synchronises). */
int modem_gpu_loopback(modem_ctx_t* ctx, const uint8_t* bits, size_t F, size_t nbits,
                       float sigma, uint64_t seed, uint64_t frame0, modem_c32_t* tx,
                       uint8_t* sym, uint8_t* bits_out, uint64_t counters[2]);

/* loopback_packed (extension; NOT the reference's payload format): modem_gpu_loopback on PACKED payloads.
 *   packed     [F][ceil(nbits/8)]   bit j of a frame = bit 7 - j%8 of byte j/8 (first bit = most significant, the order of
 *                                   digital/util.rs:5-11); pad bits of a row's last byte are ignored
 *   packed_out [F][ceil(K*bps/8)]   (nullable) the demapped bits in the same packing, pad bits zero
 * The reference's bit sources carry one BYTE per bit (data.rs:35-40, AsciiBits of modulate.rs:98); with host buffers that
 * format makes modem_gpu_loopback a PCIe copy.  This entry moves 1/8 of the bytes: the rows are
 * unpacked / packed on the device either side of the same kernels, so decisions and counters equal modem_gpu_loopback's
 * on the unpacked bits.  Host or device pointers (host: chunked copy-in / kernels / copy-out pipeline,
 * MODEM_GPU_PACKED_CHUNK frames per chunk); counters is a host pointer, the call synchronises. */
int modem_gpu_loopback_packed(modem_ctx_t* ctx, const uint8_t* packed, size_t F, size_t nbits,
                              float sigma, uint64_t seed, uint64_t frame0, uint8_t* packed_out,
                              uint64_t counters[2]);

/* ------------------------------------------------------------------ memory helpers */
int modem_gpu_malloc(modem_ctx_t* ctx, void** dptr, size_t bytes);
int modem_gpu_free(modem_ctx_t* ctx, void* dptr);
int modem_gpu_host_alloc(void** hptr, size_t bytes); /* pinned */
int modem_gpu_host_free(void* hptr);
int modem_gpu_memcpy_h2d(modem_ctx_t* ctx, void* dst, const void* src, size_t bytes);
int modem_gpu_memcpy_d2h(modem_ctx_t* ctx, void* dst, const void* src, size_t bytes);

/* ------------------------------------------------------------------ multi-GPU
 * Frames (or channels) are sharded across ranks by the caller; the only exchange is one
 * all-reduce of the error counters.  NCCL is loaded lazily (dlopen). */
#define MODEM_COMM_ID_BYTES 128
int modem_gpu_comm_unique_id(uint8_t id[MODEM_COMM_ID_BYTES]);
int modem_gpu_comm_create(modem_comm_t** comm, modem_ctx_t* ctx, int n_ranks, int rank,
                          const uint8_t id[MODEM_COMM_ID_BYTES]);
int modem_gpu_allreduce_counters(modem_comm_t* comm, uint64_t* counters, size_t n);
void modem_gpu_comm_destroy(modem_comm_t* comm);

/* ------------------------------------------------------------------ diagnostics */
const char* modem_gpu_strerror(int code);
const char* modem_gpu_last_error(const modem_ctx_t* ctx); /* detail of the last failure */
uint64_t modem_gpu_launch_count(const modem_ctx_t* ctx);  /* kernels launched by this ctx so far */
int modem_gpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MODEM_GPU_H */
