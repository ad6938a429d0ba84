//! `modem::gpu` -- thin FFI module that puts the B200 library behind the crate's public API.
//!
//! SOURCE ONLY: this image has no Rust toolchain, so this file is not compiled or tested here;
//! every symbol it binds is declared in include/modem_gpu.h and is exercised through the same
//! C ABI by the C++ mirror (rust-modem_b200/host/modem.hpp) and by tests/.  Written for a
//! current stable rustc; the reference itself needs a 2016 nightly (src/modem/lib.rs:1).
//!
//! Add `pub mod gpu;` to src/modem/lib.rs.  The streaming types keep their signatures:
//! `DigitalModulator::next` drains its `Source` on the first call, makes ONE FFI call and then
//! yields `IQSample`s from the returned buffer (see `modulate_stream`); `Demodulator::next`
//! likewise collects `sig`, calls `demodulate_stream` and yields `(f32, f32)`.
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone, Debug, Default)]
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

pub enum modem_ctx_t {}

extern "C" {
    pub fn modem_sample_freq(hz: usize, sr: usize) -> f32;
    pub fn modem_samples_per_symbol(baud_rate: usize, sample_rate: usize) -> usize;
    pub fn modem_const_by_name(name: *const c_char, out_iq: *mut f32, n_tables: *mut u32, evenodd: *mut u32) -> c_int;
    pub fn modem_const_qpsk(phase: f32, amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_qam(bps: u32, phase: f32, amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_const_mpsk(bps: u32, phase_offset: f32, amplitude: f32, out_iq: *mut f32) -> c_int;
    pub fn modem_lowpass_taps(n: *mut usize) -> *const f32;
    pub fn modem_gpu_create(ctx: *mut *mut modem_ctx_t, device: c_int, cfg: *const modem_cfg_t) -> c_int;
    pub fn modem_gpu_destroy(ctx: *mut modem_ctx_t);
    pub fn modem_gpu_frame_samples(ctx: *const modem_ctx_t, nbits: usize) -> usize;
    pub fn modem_gpu_decided_symbols(ctx: *const modem_ctx_t, l: usize) -> usize;
    pub fn modem_gpu_modulate(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize,
                              tx: *mut modem_c32_t, iq: *mut modem_c32_t) -> c_int;
    pub fn modem_gpu_awgn(ctx: *mut modem_ctx_t, buf: *mut modem_c32_t, f: usize, l: usize, sigma: f32,
                          seed: u64, frame0: u64) -> c_int;
    pub fn modem_gpu_demodulate(ctx: *mut modem_ctx_t, rx: *const modem_c32_t, f: usize, l: usize, sym: *mut u8,
                                bits: *mut u8, soft: *mut modem_c32_t, filt: *mut modem_c32_t, sigma: f32,
                                seed: u64, frame0: u64) -> c_int;
    pub fn modem_gpu_loopback(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize, sigma: f32, seed: u64,
                              frame0: u64, tx: *mut modem_c32_t, sym: *mut u8, bits_out: *mut u8,
                              counters: *mut u64) -> c_int;
    pub fn modem_hilbert_taps(n: *mut usize) -> *const f32;
    pub fn modem_phasor_by_name(name: *const c_char, baud_rate: usize, sample_rate: usize, out: *mut modem_phasor_t,
                                evenodd: *mut u32) -> c_int;
    pub fn modem_gpu_set_phasor(ctx: *mut modem_ctx_t, phasor: *const modem_phasor_t) -> c_int;
    pub fn modem_gpu_preamble(ctx: *mut modem_ctx_t, f: usize, n: usize, amplitude: f32, tx: *mut modem_c32_t) -> c_int;
    pub fn modem_gpu_modulate_real(ctx: *mut modem_ctx_t, bits: *const u8, f: usize, nbits: usize, preamble: usize,
                                   preamble_amplitude: f32, out: *mut f32) -> c_int;
    pub fn modem_gpu_lock_phase(ctx: *mut modem_ctx_t, samples: *const c_void, fmt: u32, f: usize, l: usize,
                                hilbert_taps: *const f32, n_hilbert: usize, lock_samples: usize,
                                phase_offset: *mut f32) -> c_int;
    pub fn modem_gpu_demodulate_real(ctx: *mut modem_ctx_t, samples: *const c_void, fmt: u32, f: usize, l: usize,
                                     lock_samples: usize, hilbert_taps: *const f32, n_hilbert: usize,
                                     phase_offset: *mut f32, sym: *mut u8, bits: *mut u8, soft: *mut modem_c32_t,
                                     filt: *mut modem_c32_t) -> c_int;
    pub fn modem_gpu_strerror(code: c_int) -> *const c_char;
    pub fn modem_gpu_last_error(ctx: *const modem_ctx_t) -> *const c_char;
}

/// The reference reports errors by panicking (`unwrap`/`expect`/`assert!`); so does this module.
fn check(rc: c_int, ctx: *const modem_ctx_t, what: &str) {
    if rc != 0 {
        let (a, b) = unsafe { (CStr::from_ptr(modem_gpu_strerror(rc)), CStr::from_ptr(modem_gpu_last_error(ctx))) };
        panic!("{}: {} ({})", what, a.to_string_lossy(), b.to_string_lossy());
    }
}

/// One configured path on one GPU.  Not `Sync`: one context = one device + one stream.
pub struct Context {
    raw: *mut modem_ctx_t,
    _table: Vec<f32>,
    _tx_taps: Vec<f32>,
    _rx_taps: Vec<f32>,
}

pub struct PathConfig<'a> {
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
}

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
            flags: 0,
        };
        let mut raw: *mut modem_ctx_t = std::ptr::null_mut();
        check(unsafe { modem_gpu_create(&mut raw, 0, &cfg) }, std::ptr::null(), "modem_gpu_create");
        Context { raw, _table: table, _tx_taps: tx, _rx_taps: rx }
    }

    /// What `DigitalModulator::new(..).map(|s| s.modulate())` yields for `bits` (one frame), plus the
    /// baseband `(i, q)` pairs the `--iq` flag writes (modulate.rs:109-116).
    pub fn modulate_stream(&mut self, bits: &[u8]) -> (Vec<modem_c32_t>, Vec<modem_c32_t>) {
        let l = unsafe { modem_gpu_frame_samples(self.raw, bits.len()) };
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
        let l = unsafe { modem_gpu_frame_samples(self.raw, bits.len()) };
        let mut out = vec![0f32; preamble + l];
        check(unsafe { modem_gpu_modulate_real(self.raw, bits.as_ptr(), 1, bits.len(), preamble, amplitude, out.as_mut_ptr()) },
              self.raw, "modem_gpu_modulate_real");
        out
    }

    /// `Demodulator::lock_phase` (demodulator.rs:32-36) on an analytic signal: returns `pll.phase_offset`.
    pub fn lock_phase(&mut self, sig: &[modem_c32_t]) -> f32 {
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

    /// Batched loopback over `frames` frames of `nbits` bits each: returns (bit errors, bits compared).
    pub fn loopback(&mut self, bits: &[u8], frames: usize, nbits: usize, bits_out: &mut [u8], sigma: f32, seed: u64) -> (u64, u64) {
        assert!(bits.len() >= frames * nbits);
        let mut cnt = [0u64; 2];
        check(unsafe { modem_gpu_loopback(self.raw, bits.as_ptr(), frames, nbits, sigma, seed, 0, std::ptr::null_mut(),
                                          std::ptr::null_mut(), bits_out.as_mut_ptr(), cnt.as_mut_ptr()) },
              self.raw, "modem_gpu_loopback");
        (cnt[0], cnt[1])
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { modem_gpu_destroy(self.raw) }
    }
}

#[allow(dead_code)]
fn _unused(_: *mut c_void) {}
