//! `modem::gpu` -- thin FFI module that puts the B200 library behind the crate's public API.
//!
//! SOURCE ONLY: this image has no Rust toolchain, so this file is not compiled here.  What IS checked here
//! (tests/test_rust_boundary.py): every function include/modem_gpu.h declares is bound below under the same name
//! with the same number of arguments, and build.rs compiles the same translation units with the same defines as
//! csrc/Makefile.  The same symbols are exercised through the C ABI by the C++ mirror
//! (rust-modem_b200/host/modem.hpp) and by tests/.  Written for a current stable rustc; the reference itself needs
//! a 2016 nightly (src/modem/lib.rs:1).
//!
//! Add `pub mod gpu;` to src/modem/lib.rs.  The streaming types keep their signatures:
//! `DigitalModulator::next` drains its `Source` on the first call, makes ONE FFI call and then yields `IQSample`s
//! from the returned buffer (see `modulate_stream`); `Demodulator::next` likewise collects `sig`, calls
//! `demodulate_stream` and yields `(f32, f32)`.  The call shapes of the two binaries are in INTEGRATION.md.
#![allow(non_camel_case_types)]

use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone, Debug, Default, PartialEq)]
pub struct modem_c32_t {
    pub re: f32,
    pub im: f32,
}

#[repr(C)]
pub struct modem_cfg_t {
    pub struct_size: u32,
    pub bits_per_symbol: u32,
    pub samples_per_symbol: u32,
    pub n_tables: u32,
    pub const_iq: *const f32,
    pub q_offset: u32,
    pub sample_freq: f32,
    pub phase_offset: f32,
    pub sample0: u64,
    pub n_tx_taps: u32,
    pub tx_taps: *const f32,
    pub n_rx_taps: u32,
    pub rx_taps: *const f32,
    pub decision_delay: u32,
    pub rx_gain: f32,
    pub slicer_gain: f32,
    pub flags: u32,
}
pub const MODEM_FLAG_FUSED_MAC: u32 = 0x1;
pub const MODEM_FLAG_NO_TMEM: u32 = 0x2;

/// apsk.rs:60-67
#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct modem_ring_t {
    pub start: u8,
    pub end: u8,
    pub radius: f32,
    pub phase: f32,
}

/// Stateful / time-varying mappers (digital/{bfsk,mfsk,cpfsk,msk,dmpsk}.rs): `modem_phasor_t`.
#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
pub struct modem_phasor_t {
    pub struct_size: u32,
    pub kind: u32, // MODEM_PHASOR_*: 1 bfsk, 2 mfsk, 3 cpfsk, 4 msk, 5 dmpsk
    pub bits_per_symbol: u32,
    pub amplitude: f32,
    pub deviation: f32,
    pub phase: f32,
    pub shift: f32,
    pub mfsk_increase_map: u32,
}
pub const MODEM_SAMPLES_C32: u32 = 0;
pub const MODEM_SAMPLES_F32: u32 = 1;
pub const MODEM_SAMPLES_I16: u32 = 2;
pub const LOCK_SAMPLES: usize = 64; // demodulator.rs:5
pub const MODEM_COMM_ID_BYTES: usize = 128;

pub enum modem_ctx_t {}
pub enum modem_comm_t {}

extern "C" {
    // ---- host-side helpers (no device needed)
    pub fn modem_sample_freq(hz: usize, sr: usize) -> f32;
    pub fn modem_samples_per_symbol(baud_rate: usize, sample_rate: usize) -> usize;
    pub fn modem_const_bask(amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_bpsk(phase: f32, amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_qpsk(phase: f32, amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_qam(bps: u32, phase: f32, amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_mpsk(bps: u32, phase_offset: f32, amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_oqpsk(amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_dcqpsk(amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_apsk(amplitude: f32, bps: u32, rings: *const modem_ring_t, n_rings: usize, out_iq: *mut f32) -> c_int;
    pub fn modem_const_by_name(name: *const c_char, out_iq: *mut f32, n_tables: *mut u32, evenodd: *mut u32) -> c_int;
    pub fn modem_lowpass_taps(n: *mut usize) -> *const f32;
    pub fn modem_hilbert_taps(n: *mut usize) -> *const f32;
    pub fn modem_phasor_by_name(name: *const c_char, baud_rate: usize, sample_rate: usize, out: *mut modem_phasor_t,
                                evenodd: *mut u32) -> c_int;
    pub fn modem_rrc_taps(out: *mut f32, span: usize, sps: usize, beta: c_double) -> c_int;
    pub fn modem_sigma_for_ebn0(cfg: *const modem_cfg_t, ebn0_db: c_double) -> f32;
    // ---- context
    pub fn modem_gpu_device_count(n: *mut c_int) -> c_int;
    pub fn modem_gpu_create(ctx: *mut *mut modem_ctx_t, device: c_int, cfg: *const modem_cfg_t) -> c_int;
    pub fn modem_gpu_destroy(ctx: *mut modem_ctx_t);
    pub fn modem_gpu_set_stream(ctx: *mut modem_ctx_t, cuda_stream: *mut c_void) -> c_int;
    pub fn modem_gpu_set_channels(ctx: *mut modem_ctx_t, n_channels: usize, sample_freq: *const f32,
                                  phase_offset: *const f32, frames_per_channel: usize) -> c_int;
    pub fn modem_gpu_synchronize(ctx: *mut modem_ctx_t) -> c_int;
    pub fn modem_gpu_set_phasor(ctx: *mut modem_ctx_t, phasor: *const modem_phasor_t) -> c_int;
    pub fn modem_gpu_frame_samples(ctx: *const modem_ctx_t, nbits: usize) -> usize;
    pub fn modem_gpu_decided_symbols(ctx: *const modem_ctx_t, l: usize) -> usize;
    // ---- the path
    pub fn modem_gpu_modulate(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize,
                              tx: *mut modem_c32_t, iq: *mut modem_c32_t) -> c_int;
    pub fn modem_gpu_preamble(ctx: *mut modem_ctx_t, f: usize, n: usize, amplitude: f32, tx: *mut modem_c32_t) -> c_int;
    pub fn modem_gpu_modulate_real(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize, preamble: usize,
                                   preamble_amplitude: f32, out: *mut f32) -> c_int;
    pub fn modem_gpu_awgn(ctx: *mut modem_ctx_t, buf: *mut modem_c32_t, f: usize, l: usize, sigma: f32,
                          seed: u64, frame0: u64) -> c_int;
    pub fn modem_gpu_random_bits(ctx: *mut modem_ctx_t, bits: *mut u8, f: usize, nbits: usize, seed: u64, frame0: u64) -> c_int;
    pub fn modem_gpu_demodulate(ctx: *mut modem_ctx_t, rx: *const modem_c32_t, f: usize, l: usize, sym: *mut u8,
                                bits: *mut u8, soft: *mut modem_c32_t, filt: *mut modem_c32_t, sigma: f32,
                                seed: u64, frame0: u64) -> c_int;
    pub fn modem_gpu_lock_phase(ctx: *mut modem_ctx_t, samples: *const c_void, fmt: u32, f: usize, l: usize,
                                hilbert_taps: *const f32, n_hilbert: usize, lock_samples: usize,
                                phase_offset: *mut f32) -> c_int;
    pub fn modem_gpu_demodulate_real(ctx: *mut modem_ctx_t, samples: *const c_void, fmt: u32, f: usize, l: usize,
                                     lock_samples: usize, hilbert_taps: *const f32, n_hilbert: usize,
                                     phase_offset: *mut f32, sym: *mut u8, bits: *mut u8, soft: *mut modem_c32_t,
                                     filt: *mut modem_c32_t) -> c_int;
    pub fn modem_gpu_demodulate_count(ctx: *mut modem_ctx_t, rx: *const modem_c32_t, f: usize, l: usize, sym: *mut u8,
                                      bits: *mut u8, ref_bits: *const u8, ref_stride: usize, counters: *mut u64,
                                      sigma: f32, seed: u64, frame0: u64) -> c_int;
    pub fn modem_gpu_ber_sweep(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize, n_points: usize,
                               sigmas: *const f32, seed: u64, frame0: u64, tx: *mut modem_c32_t,
                               counters: *mut u64) -> c_int;
    pub fn modem_gpu_loopback_device(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize, sigma: f32,
                                     seed: u64, frame0: u64, tx: *mut modem_c32_t, sym: *mut u8, bits_out: *mut u8,
                                     counters: *mut u64) -> c_int;
    pub fn modem_gpu_loopback(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize, sigma: f32, seed: u64,
                              frame0: u64, tx: *mut modem_c32_t, sym: *mut u8, bits_out: *mut u8,
                              counters: *mut u64) -> c_int;
    pub fn modem_gpu_loopback_packed(ctx: *mut modem_ctx_t, packed: *const u8, f: usize, nbits: usize, sigma: f32,
                                     seed: u64, frame0: u64, packed_out: *mut u8, counters: *mut u64) -> c_int;
    // ---- memory helpers
    pub fn modem_gpu_malloc(ctx: *mut modem_ctx_t, dptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn modem_gpu_free(ctx: *mut modem_ctx_t, dptr: *mut c_void) -> c_int;
    pub fn modem_gpu_host_alloc(hptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn modem_gpu_host_free(hptr: *mut c_void) -> c_int;
    pub fn modem_gpu_memcpy_h2d(ctx: *mut modem_ctx_t, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn modem_gpu_memcpy_d2h(ctx: *mut modem_ctx_t, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    // ---- multi-GPU: the single exchange of the path
    pub fn modem_gpu_comm_unique_id(id: *mut u8) -> c_int;
    pub fn modem_gpu_comm_create(comm: *mut *mut modem_comm_t, ctx: *mut modem_ctx_t, n_ranks: c_int, rank: c_int,
                                 id: *const u8) -> c_int;
    pub fn modem_gpu_allreduce_counters(comm: *mut modem_comm_t, counters: *mut u64, n: usize) -> c_int;
    pub fn modem_gpu_comm_destroy(comm: *mut modem_comm_t);
    // ---- diagnostics
    pub fn modem_gpu_strerror(code: c_int) -> *const c_char;
    pub fn modem_gpu_last_error(ctx: *const modem_ctx_t) -> *const c_char;
    pub fn modem_gpu_launch_count(ctx: *const modem_ctx_t) -> u64;
    pub fn modem_gpu_abi_version() -> c_int;
}

/// The reference reports errors by panicking (`unwrap`/`expect`/`assert!`); so does this module.
fn check(rc: c_int, ctx: *const modem_ctx_t, what: &str) {
    if rc != 0 {
        let (a, b) = unsafe { (CStr::from_ptr(modem_gpu_strerror(rc)), CStr::from_ptr(modem_gpu_last_error(ctx))) };
        panic!("{}: {} ({})", what, a.to_string_lossy(), b.to_string_lossy());
    }
}

/// `(table [n_tables][2^bps][2], bps, n_tables, evenodd)` for a memoryless `-m` name of modulate.rs:74-95;
/// panics with the reference's message for an unknown name (modulate.rs:94).
pub fn constellation_by_name(name: &str) -> (Vec<f32>, usize, usize, bool) {
    let c = CString::new(name).unwrap();
    let mut out = vec![0f32; 2 * 512];
    let (mut nt, mut eo) = (0u32, 0u32);
    let bps = unsafe { modem_const_by_name(c.as_ptr(), out.as_mut_ptr(), &mut nt, &mut eo) };
    assert!(bps > 0, "invalid digital modulation");
    out.truncate(2 * (nt as usize) << bps);
    (out, bps as usize, nt as usize, eo != 0)
}

/// The 64-tap low-pass of src/bin/demodulate.rs:82-147.
pub fn lowpass_taps() -> &'static [f32] {
    let mut n = 0usize;
    unsafe {
        let p = modem_lowpass_taps(&mut n);
        std::slice::from_raw_parts(p, n)
    }
}

/// One configured path on one GPU.  Not `Sync`: one context = one device + one stream.
pub struct Context {
    raw: *mut modem_ctx_t,
    bps: usize,
    _table: Vec<f32>,
    _tx_taps: Vec<f32>,
    _rx_taps: Vec<f32>,
}

pub struct PathConfig<'a> {
    pub device: usize,
    pub bits_per_symbol: usize,
    pub n_tables: usize,
    pub table: Vec<f32>, // [n_tables][2^bps][2] from the digital::* formulas (modem_const_*)
    pub samples_per_symbol: usize,
    pub sample_freq: f32, // Freq::sample_freq()
    pub phase_offset: f32, // PLL::phase_offset
    pub sample0: usize, // Carrier.sample
    pub q_offset: usize, // EvenOddOffset: samples_per_symbol / 2, else 0
    pub tx_taps: &'a [f32], // empty => rectangular hold (the reference's TX)
    pub rx_taps: &'a [f32], // the `lp` closure's taps
    pub decision_delay: usize,
    pub slicer_gain: f32,
    pub flags: u32, // MODEM_FLAG_*
}

/// Error counters of a loopback / sweep: (bit errors, bits compared).
pub type Counters = (u64, u64);

impl Context {
    pub fn new(p: PathConfig) -> Context {
        let (table, tx, rx) = (p.table, p.tx_taps.to_vec(), p.rx_taps.to_vec());
        let cfg = modem_cfg_t {
            struct_size: std::mem::size_of::<modem_cfg_t>() as u32,
            bits_per_symbol: p.bits_per_symbol as u32,
            samples_per_symbol: p.samples_per_symbol as u32,
            n_tables: p.n_tables as u32,
            const_iq: table.as_ptr(),
            q_offset: p.q_offset as u32,
            sample_freq: p.sample_freq,
            phase_offset: p.phase_offset,
            sample0: p.sample0 as u64,
            n_tx_taps: tx.len() as u32,
            tx_taps: if tx.is_empty() { std::ptr::null() } else { tx.as_ptr() },
            n_rx_taps: rx.len() as u32,
            rx_taps: rx.as_ptr(),
            decision_delay: p.decision_delay as u32,
            rx_gain: 2.0, // demodulator.rs:53-54
            slicer_gain: p.slicer_gain,
            flags: p.flags,
        };
        let mut raw: *mut modem_ctx_t = std::ptr::null_mut();
        check(unsafe { modem_gpu_create(&mut raw, p.device as c_int, &cfg) }, std::ptr::null(), "modem_gpu_create");
        Context { raw, bps: p.bits_per_symbol, _table: table, _tx_taps: tx, _rx_taps: rx }
    }

    pub fn frame_samples(&self, nbits: usize) -> usize {
        unsafe { modem_gpu_frame_samples(self.raw, nbits) }
    }
    pub fn decided_symbols(&self, l: usize) -> usize {
        unsafe { modem_gpu_decided_symbols(self.raw, l) }
    }
    /// Kernels launched by this context so far (a caller's proof that the GPU did the work).
    pub fn launch_count(&self) -> u64 {
        unsafe { modem_gpu_launch_count(self.raw) }
    }
    pub fn synchronize(&mut self) {
        check(unsafe { modem_gpu_synchronize(self.raw) }, self.raw, "modem_gpu_synchronize");
    }
    /// Run this context's work on the caller's CUDA stream (`cudaStream_t` / `CUstream`).
    ///
    /// # Safety
    /// `cuda_stream` must be a live stream of this context's device (or null for the legacy default stream).
    pub unsafe fn set_stream(&mut self, cuda_stream: *mut c_void) {
        check(modem_gpu_set_stream(self.raw, cuda_stream), self.raw, "modem_gpu_set_stream");
    }
    /// Multi-carrier bank: frame `f` uses `sample_freq[f / frames_per_channel]` (and `phase_offset[..]` when given).
    pub fn set_channels(&mut self, sample_freq: &[f32], phase_offset: Option<&[f32]>, frames_per_channel: usize) {
        if let Some(po) = phase_offset {
            assert_eq!(po.len(), sample_freq.len());
        }
        let po = phase_offset.map_or(std::ptr::null(), |p| p.as_ptr());
        check(unsafe { modem_gpu_set_channels(self.raw, sample_freq.len(), sample_freq.as_ptr(), po, frames_per_channel) },
              self.raw, "modem_gpu_set_channels");
    }

    /// What `DigitalModulator::new(..).map(|s| s.modulate())` yields for `bits` (one frame), plus the
    /// baseband `(i, q)` pairs the `--iq` flag writes (modulate.rs:109-116).
    pub fn modulate_stream(&mut self, bits: &[u8]) -> (Vec<modem_c32_t>, Vec<modem_c32_t>) {
        let l = self.frame_samples(bits.len());
        let (mut tx, mut iq) = (vec![modem_c32_t::default(); l], vec![modem_c32_t::default(); l]);
        check(unsafe { modem_gpu_modulate(self.raw, bits.as_ptr(), 1, bits.len(), tx.as_mut_ptr(), iq.as_mut_ptr()) },
              self.raw, "modem_gpu_modulate");
        (tx, iq)
    }

    /// What iterating `Demodulator` yields for `sig` (one frame): the full-rate filtered `(I, Q)` stream.
    pub fn demodulate_stream(&mut self, sig: &[modem_c32_t]) -> Vec<modem_c32_t> {
        let mut filt = vec![modem_c32_t::default(); sig.len()];
        check(unsafe { modem_gpu_demodulate(self.raw, sig.as_ptr(), 1, sig.len(), std::ptr::null_mut(), std::ptr::null_mut(),
                                            std::ptr::null_mut(), filt.as_mut_ptr(), 0.0, 0, 0) },
              self.raw, "modem_gpu_demodulate");
        filt
    }

    /// Switch the TX mapper to one of the stateful schemes (`bfsk`, `mfsk`, `16cpfsk`, `msk`, `dqpsk`, `dbpsk`:
    /// modulate.rs:74-95); the `DigitalPhasor::update` recurrence then runs in the phasor kernels.
    pub fn set_phasor(&mut self, p: &modem_phasor_t) {
        check(unsafe { modem_gpu_set_phasor(self.raw, p) }, self.raw, "modem_gpu_set_phasor");
    }

    /// `Modulator::new(&mut carrier, Box::new(phasor::Raw::new(amplitude))).take(n)` mapped through
    /// `modulate()` (modulate.rs:118-126).
    pub fn preamble(&mut self, n: usize, amplitude: f32) -> Vec<modem_c32_t> {
        let mut tx = vec![modem_c32_t::default(); n];
        check(unsafe { modem_gpu_preamble(self.raw, 1, n, amplitude, tx.as_mut_ptr()) }, self.raw, "modem_gpu_preamble");
        tx
    }

    /// Everything src/bin/modulate.rs writes without `--iq` for one stream: sync tone, then data, real part.
    pub fn modulate_real(&mut self, bits: &[u8], preamble: usize, amplitude: f32) -> Vec<f32> {
        let l = self.frame_samples(bits.len());
        let mut out = vec![0f32; preamble + l];
        check(unsafe { modem_gpu_modulate_real(self.raw, bits.as_ptr(), 1, bits.len(), preamble, amplitude, out.as_mut_ptr()) },
              self.raw, "modem_gpu_modulate_real");
        out
    }

    /// `Demodulator::lock_phase` (demodulator.rs:32-36) on an analytic signal: returns `pll.phase_offset`.
    pub fn lock_phase(&mut self, sig: &[modem_c32_t]) -> f32 {
        assert!(sig.len() >= LOCK_SAMPLES, "called `Option::unwrap()` on a `None` value"); // demodulator.rs:34
        let mut po = 0f32;
        check(unsafe { modem_gpu_lock_phase(self.raw, sig.as_ptr() as *const c_void, MODEM_SAMPLES_C32, 1, sig.len(),
                                            std::ptr::null(), 0, LOCK_SAMPLES, &mut po) },
              self.raw, "modem_gpu_lock_phase");
        po
    }

    /// src/bin/demodulate.rs:29-43 for one stream of native-endian i16 samples: Hilbert FIR + 64-sample PLL lock +
    /// the two low-pass FIRs; returns (locked phase offset, the `(i, q)` stream the binary prints).
    pub fn demodulate_i16(&mut self, samples: &[i16]) -> (f32, Vec<modem_c32_t>) {
        assert!(samples.len() >= LOCK_SAMPLES, "called `Option::unwrap()` on a `None` value"); // demodulator.rs:34
        let mut po = 0f32;
        let mut filt = vec![modem_c32_t::default(); samples.len() - LOCK_SAMPLES];
        check(unsafe { modem_gpu_demodulate_real(self.raw, samples.as_ptr() as *const c_void, MODEM_SAMPLES_I16, 1,
                                                 samples.len(), LOCK_SAMPLES, std::ptr::null(), 0, &mut po,
                                                 std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut(),
                                                 filt.as_mut_ptr()) },
              self.raw, "modem_gpu_demodulate_real");
        (po, filt)
    }

    /// In-place AWGN on `frames` frames of `l` samples (extension; oracle/modem_oracle.h "AWGN").
    pub fn awgn(&mut self, buf: &mut [modem_c32_t], frames: usize, l: usize, sigma: f32, seed: u64, frame0: u64) {
        assert!(buf.len() >= frames * l);
        check(unsafe { modem_gpu_awgn(self.raw, buf.as_mut_ptr(), frames, l, sigma, seed, frame0) }, self.raw, "modem_gpu_awgn");
    }

    /// Philox payload bits for `frames` frames of `nbits` bits (one byte per bit).
    pub fn random_bits(&mut self, frames: usize, nbits: usize, seed: u64, frame0: u64) -> Vec<u8> {
        let mut bits = vec![0u8; frames * nbits];
        check(unsafe { modem_gpu_random_bits(self.raw, bits.as_mut_ptr(), frames, nbits, seed, frame0) }, self.raw, "modem_gpu_random_bits");
        bits
    }

    /// Batched loopback over `frames` frames of `nbits` bits each with HOST buffers: modulate -> (AWGN) -> demodulate ->
    /// count errors against the input bits.  `bits_out` receives the demapped bits, `frames * K * bps` bytes with
    /// `K = decided_symbols(frame_samples(nbits))`.
    pub fn loopback(&mut self, bits: &[u8], frames: usize, nbits: usize, bits_out: &mut [u8], sigma: f32, seed: u64) -> Counters {
        assert!(bits.len() >= frames * nbits);
        let k = self.decided_symbols(self.frame_samples(nbits));
        assert!(bits_out.len() >= frames * k * self.bps, "bits_out holds {} bytes, the call writes {}", bits_out.len(), frames * k * self.bps);
        let mut cnt = [0u64; 2];
        check(unsafe { modem_gpu_loopback(self.raw, bits.as_ptr(), frames, nbits, sigma, seed, 0, std::ptr::null_mut(),
                                          std::ptr::null_mut(), bits_out.as_mut_ptr(), cnt.as_mut_ptr()) },
              self.raw, "modem_gpu_loopback");
        (cnt[0], cnt[1])
    }

    /// Extension (not the reference's payload format): the loopback on PACKED payloads, 8 bits per byte with the first bit
    /// in the most significant position (the order of `digital::util::bytes_to_bits`).  `packed` holds `frames` rows of
    /// `ceil(nbits / 8)` bytes, `packed_out` receives `frames` rows of `ceil(K * bps / 8)` bytes.  Same decisions and
    /// counters as `loopback` on the unpacked bits, 1/8 of the bytes over PCIe.
    pub fn loopback_packed(&mut self, packed: &[u8], frames: usize, nbits: usize, packed_out: &mut [u8], sigma: f32, seed: u64) -> Counters {
        assert!(packed.len() >= frames * ((nbits + 7) / 8));
        let k = self.decided_symbols(self.frame_samples(nbits));
        let ob = (k * self.bps + 7) / 8;
        assert!(packed_out.len() >= frames * ob, "packed_out holds {} bytes, the call writes {}", packed_out.len(), frames * ob);
        let mut cnt = [0u64; 2];
        check(unsafe { modem_gpu_loopback_packed(self.raw, packed.as_ptr(), frames, nbits, sigma, seed, 0, packed_out.as_mut_ptr(),
                                                 cnt.as_mut_ptr()) },
              self.raw, "modem_gpu_loopback_packed");
        (cnt[0], cnt[1])
    }

    /// Device memory owned by this context's device (freed by `DeviceBuffer::drop`).
    pub fn device_alloc(&mut self, bytes: usize) -> DeviceBuffer {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { modem_gpu_malloc(self.raw, &mut p, bytes) }, self.raw, "modem_gpu_malloc");
        DeviceBuffer { ctx: self.raw, ptr: p, bytes }
    }
    pub fn upload(&mut self, dst: &mut DeviceBuffer, src: &[u8]) {
        assert!(src.len() <= dst.bytes);
        check(unsafe { modem_gpu_memcpy_h2d(self.raw, dst.ptr, src.as_ptr() as *const c_void, src.len()) }, self.raw, "modem_gpu_memcpy_h2d");
    }
    pub fn download(&mut self, dst: &mut [u8], src: &DeviceBuffer) {
        assert!(dst.len() <= src.bytes);
        check(unsafe { modem_gpu_memcpy_d2h(self.raw, dst.as_mut_ptr() as *mut c_void, src.ptr, dst.len()) }, self.raw, "modem_gpu_memcpy_d2h");
    }

    /// The whole loopback, device-resident and stream-ordered: for the headline shape ONE fused kernel that makes the TX
    /// samples, stores them to `tx` (not at all when `None`) and demodulates them.  `counters` is a device `u64[2]` that is
    /// accumulated into.  Every buffer is checked against the size the call writes.
    pub fn loopback_device(&mut self, bits: &DeviceBuffer, frames: usize, nbits: usize, sigma: f32, seed: u64, frame0: u64,
                           tx: Option<&mut DeviceBuffer>, sym: Option<&mut DeviceBuffer>, bits_out: Option<&mut DeviceBuffer>,
                           counters: &mut DeviceBuffer) {
        let l = self.frame_samples(nbits);
        let k = self.decided_symbols(l);
        assert!(bits.bytes >= frames * nbits && counters.bytes >= 16);
        let tx_p = tx.map_or(std::ptr::null_mut(), |b| { assert!(b.bytes >= frames * l * 8); b.ptr as *mut modem_c32_t });
        let sym_p = sym.map_or(std::ptr::null_mut(), |b| { assert!(b.bytes >= frames * k); b.ptr as *mut u8 });
        let out_p = bits_out.map_or(std::ptr::null_mut(), |b| { assert!(b.bytes >= frames * k * self.bps); b.ptr as *mut u8 });
        check(unsafe { modem_gpu_loopback_device(self.raw, bits.ptr as *const u8, frames, nbits, sigma, seed, frame0, tx_p, sym_p, out_p,
                                                 counters.ptr as *mut u64) },
              self.raw, "modem_gpu_loopback_device");
    }

    /// Demodulate device-resident samples and ACCUMULATE bit errors against `ref_bits` into the device counters.
    pub fn demodulate_count(&mut self, rx: &DeviceBuffer, frames: usize, l: usize, ref_bits: &DeviceBuffer, ref_stride: usize,
                            counters: &mut DeviceBuffer, sigma: f32, seed: u64, frame0: u64) {
        assert!(rx.bytes >= frames * l * 8 && ref_bits.bytes >= frames * ref_stride && counters.bytes >= 16);
        check(unsafe { modem_gpu_demodulate_count(self.raw, rx.ptr as *const modem_c32_t, frames, l, std::ptr::null_mut(), std::ptr::null_mut(),
                                                  ref_bits.ptr as *const u8, ref_stride, counters.ptr as *mut u64, sigma, seed, frame0) },
              self.raw, "modem_gpu_demodulate_count");
    }

    /// BASELINE config 4: modulate the frames once, then one noisy demodulation per entry of `sigmas`; the device array
    /// `counters[p][2]` is accumulated into.  Shard frames across ranks with `frame0`, reduce with `Comm::allreduce`.
    pub fn ber_sweep(&mut self, bits: &DeviceBuffer, frames: usize, nbits: usize, sigmas: &[f32], seed: u64, frame0: u64,
                     counters: &mut DeviceBuffer) {
        assert!(bits.bytes >= frames * nbits && counters.bytes >= 16 * sigmas.len());
        check(unsafe { modem_gpu_ber_sweep(self.raw, bits.ptr as *const u8, frames, nbits, sigmas.len(), sigmas.as_ptr(), seed, frame0,
                                           std::ptr::null_mut(), counters.ptr as *mut u64) },
              self.raw, "modem_gpu_ber_sweep");
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { modem_gpu_destroy(self.raw) }
    }
}

/// Device memory of one context (`modem_gpu_malloc` / `modem_gpu_free`).
pub struct DeviceBuffer {
    ctx: *mut modem_ctx_t,
    ptr: *mut c_void,
    bytes: usize,
}
impl DeviceBuffer {
    pub fn len(&self) -> usize {
        self.bytes
    }
    pub fn as_ptr(&self) -> *mut c_void {
        self.ptr
    }
}
impl Drop for DeviceBuffer {
    fn drop(&mut self) {
        unsafe { modem_gpu_free(self.ctx, self.ptr) };
    }
}

/// Pinned host memory (`modem_gpu_host_alloc`): what the host-buffer loopback copies from / to at full PCIe speed.
pub struct PinnedBuffer {
    ptr: *mut c_void,
    bytes: usize,
}
impl PinnedBuffer {
    pub fn new(bytes: usize) -> PinnedBuffer {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { modem_gpu_host_alloc(&mut p, bytes) }, std::ptr::null(), "modem_gpu_host_alloc");
        PinnedBuffer { ptr: p, bytes }
    }
    pub fn as_slice(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.ptr as *const u8, self.bytes) }
    }
    pub fn as_mut_slice(&mut self) -> &mut [u8] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr as *mut u8, self.bytes) }
    }
}
impl Drop for PinnedBuffer {
    fn drop(&mut self) {
        unsafe { modem_gpu_host_free(self.ptr) };
    }
}

/// The single exchange of the multi-GPU path: one all-reduce of the error counters.
pub struct Comm {
    raw: *mut modem_comm_t,
}
impl Comm {
    /// Rank 0 creates the id and hands it to the other ranks (any transport: it is 128 opaque bytes).
    pub fn unique_id() -> [u8; MODEM_COMM_ID_BYTES] {
        let mut id = [0u8; MODEM_COMM_ID_BYTES];
        check(unsafe { modem_gpu_comm_unique_id(id.as_mut_ptr()) }, std::ptr::null(), "modem_gpu_comm_unique_id");
        id
    }
    pub fn new(ctx: &mut Context, n_ranks: usize, rank: usize, id: &[u8; MODEM_COMM_ID_BYTES]) -> Comm {
        let mut raw: *mut modem_comm_t = std::ptr::null_mut();
        check(unsafe { modem_gpu_comm_create(&mut raw, ctx.raw, n_ranks as c_int, rank as c_int, id.as_ptr()) }, ctx.raw, "modem_gpu_comm_create");
        Comm { raw }
    }
    /// Sum `counters` (host memory) over all ranks, in place.
    pub fn allreduce(&mut self, counters: &mut [u64]) {
        check(unsafe { modem_gpu_allreduce_counters(self.raw, counters.as_mut_ptr(), counters.len()) }, std::ptr::null(), "modem_gpu_allreduce_counters");
    }
    /// Sum device-resident counters over all ranks, in place, on the context's stream (nothing is synchronised).
    pub fn allreduce_device(&mut self, counters: &mut DeviceBuffer, n: usize) {
        assert!(counters.bytes >= 8 * n);
        check(unsafe { modem_gpu_allreduce_counters(self.raw, counters.ptr as *mut u64, n) }, std::ptr::null(), "modem_gpu_allreduce_counters");
    }
}
impl Drop for Comm {
    fn drop(&mut self) {
        unsafe { modem_gpu_comm_destroy(self.raw) }
    }
}

pub fn device_count() -> usize {
    let mut n: c_int = 0;
    check(unsafe { modem_gpu_device_count(&mut n) }, std::ptr::null(), "modem_gpu_device_count");
    n as usize
}
