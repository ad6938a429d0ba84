/*
 * modem.hpp -- C++ mirror of the reference's public API (crate `modem`, /root/reference/src/modem)
 * on top of the C ABI of include/modem_gpu.h.
 *
 * The reference is Rust and this image has no Rust toolchain, so the host side that a src/bin
 * caller sees is restated here in C++ with the SAME module / type / method names, argument
 * meaning and error behaviour (Rust panics / assert! become modem::Panic exceptions).  The
 * Rust shim that binds the same symbols is in rust-modem_b200/rust/ (source only).
 *
 *   modem::freq::Freq            freq.rs          modem::rates::Rates          rates.rs
 *   modem::carrier::Carrier      carrier.rs       modem::fir::FIRFilter        fir.rs
 *   modem::data::{Source, SourceUpdate, Bits, EvenOddOffset}               data.rs
 *   modem::digital::{DigitalPhasor, bask::BASK, bpsk::BPSK, qpsk::QPSK, qam::QAM, mpsk::MPSK,
 *                    oqpsk::OQPSK, dcqpsk::DCQPSK, apsk::{APSK, Ring}}   digital/<scheme>.rs
 *   modem::digital::{bfsk::BFSK, mfsk::{MFSK, DefaultMap, IncreaseMap}, cpfsk::CPFSK, msk::MSK,
 *                    dmpsk::DMPSK}                                       the stateful schemes (TX)
 *   modem::data::AsciiBits                                                 data.rs:125-186
 *   modem::phasor::{Phasor, Raw}, modem::modulator::Modulator              phasor.rs, modulator.rs:8-62
 *   modem::modulator::{IQSample, DigitalModulator}                        modulator.rs
 *   modem::demodulator::Demodulator                                        demodulator.rs
 *
 * What differs, by design: the reference pulls ONE sample per Iterator::next() through three
 * virtual calls; here the first next() drains the Source (integer work, data.rs semantics kept
 * on the host), makes ONE call into the CUDA library for the whole stream, and the following
 * next() calls hand out the precomputed samples.  modem::gpu::{ModulatorBatch, DemodulatorBatch,
 * Loopback} expose the batched [frames][samples] form directly.
 *
 * No CPU fallback: sample arithmetic happens only in the CUDA library; without a GPU the first
 * next() throws modem::Panic carrying modem_gpu_strerror().
 */
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <charconv>
#include <istream>
#include <iterator>
#include <type_traits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/modem_gpu.h"

namespace modem {

/* Rust `panic!` / `assert!` / `unwrap()` on the reference side. */
struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc, modem_ctx_t* ctx, const char* what)
{
    if (rc != MODEM_OK) {
        const char* detail = modem_gpu_last_error(ctx);
        throw Panic(std::string(what) + ": " + modem_gpu_strerror(rc) + (detail && *detail ? std::string(" (") + detail + ")" : ""));
    }
}

using Complex32 = modem_c32_t; /* num::Complex<f32> */

/* ------------------------------------------------------------------ freq.rs, rates.rs */
namespace freq {
struct Freq {
    size_t hz, sr;
    Freq(size_t hz_, size_t sr_) : hz(hz_), sr(sr_) {}                       /* freq.rs:11-16 */
    float sample_freq() const { return modem_sample_freq(hz, sr); }           /* freq.rs:24-26 */
};
} // namespace freq

namespace rates {
struct Rates {
    size_t baud_rate, sample_rate, samples_per_symbol;
    Rates(size_t br, size_t sr) : baud_rate(br), sample_rate(sr), samples_per_symbol(modem_samples_per_symbol(br, sr)) {} /* rates.rs:12-18 */
};
} // namespace rates

/* ------------------------------------------------------------------ carrier.rs */
namespace carrier {
struct Carrier {
    float sample_freq;
    size_t sample; /* pub field (carrier.rs:6): carries over from a preamble to the data (modulate.rs:120,128) */
    explicit Carrier(const freq::Freq& f) : sample_freq(f.sample_freq()), sample(0) {} /* carrier.rs:10-15 */
};
} // namespace carrier

/* ------------------------------------------------------------------ fir.rs */
namespace fir {
/* Borrows its taps like FIRFilter<'a> (fir.rs:3-7).  On this path it is a description of the
 * filter handed to Demodulator::new; the convolution itself runs in the RX kernels. */
struct FIRFilter {
    const float* coefs;
    size_t len;
    FIRFilter(const float* c, size_t n) : coefs(c), len(n) {} /* fir.rs:10-16 */
};
/* src/bin/demodulate.rs:81-150 */
inline FIRFilter lowpass()
{
    size_t n = 0;
    const float* p = modem_lowpass_taps(&n);
    return FIRFilter(p, n);
}
/* src/bin/demodulate.rs:47-75 */
inline FIRFilter hilbert()
{
    size_t n = 0;
    const float* p = modem_hilbert_taps(&n);
    return FIRFilter(p, n);
}
} // namespace fir

/* ------------------------------------------------------------------ data.rs */
namespace data {
struct SourceUpdate { /* data.rs:4-8 */
    enum Kind { Changed, Unchanged, Finished } kind;
    const uint8_t* bits;
    size_t len;
    bool operator==(const SourceUpdate& o) const
    {
        return kind == o.kind && (kind == Finished || (len == o.len && !std::memcmp(bits, o.bits, len)));
    }
};
struct Source { /* data.rs:10-12 */
    virtual ~Source() = default;
    virtual SourceUpdate next() = 0;
    virtual size_t q_offset() const { return 0; } /* samples the odd bit lags (EvenOddOffset), else 0 */
};
class SymbolClock { /* data.rs:14-33 */
    size_t samples_per_symbol, counter;
public:
    explicit SymbolClock(size_t sps) : samples_per_symbol(sps), counter(sps - 1) {}
    bool next()
    {
        counter += 1;
        counter %= samples_per_symbol;
        return counter == 0;
    }
};
class Bits : public Source { /* data.rs:35-79 */
    const uint8_t* bits_;
    size_t nbits_;
    SymbolClock clock_;
    size_t bits_per_symbol_, idx_ = 0;
    const uint8_t* slice() const
    {
        const size_t start = (idx_ - 1) * bits_per_symbol_, end = start + bits_per_symbol_;
        return end <= nbits_ ? bits_ + start : nullptr; /* data.rs:54-63: a partial tail symbol ends the stream */
    }
public:
    Bits(const uint8_t* bits, size_t nbits, size_t samples_per_symbol, size_t bits_per_symbol)
        : bits_(bits), nbits_(nbits), clock_(samples_per_symbol), bits_per_symbol_(bits_per_symbol) {}
    SourceUpdate next() override
    {
        if (clock_.next()) {
            idx_ += 1;
            const uint8_t* b = slice();
            return b ? SourceUpdate{SourceUpdate::Changed, b, bits_per_symbol_} : SourceUpdate{SourceUpdate::Finished, nullptr, 0};
        }
        return SourceUpdate{SourceUpdate::Unchanged, slice(), bits_per_symbol_};
    }
    /* batch view used by the GPU adapters: the bit slice and the symbol count the iterator would yield */
    const uint8_t* raw() const { return bits_; }
    size_t raw_len() const { return nbits_; }
};
template <class D>
class EvenOddOffset : public Source { /* data.rs:81-123 */
    D data_;
    SymbolClock clock_;
    uint8_t cur_[2] = {0, 0};
    size_t q_off_;
public:
    EvenOddOffset(D d, size_t samples_per_symbol, size_t bits_per_symbol)
        : data_(std::move(d)), clock_(samples_per_symbol / (bits_per_symbol ? bits_per_symbol : 1)), q_off_(samples_per_symbol / 2)
    {
        if (bits_per_symbol != 2) throw Panic("assertion failed: bits_per_symbol == 2");                    /* data.rs:91 */
        if (samples_per_symbol % bits_per_symbol) throw Panic("assertion failed: samples_per_symbol % bits_per_symbol == 0"); /* :92 */
    }
    SourceUpdate next() override
    {
        SourceUpdate in = data_.next();
        if (in.kind == SourceUpdate::Finished) return in;
        if (in.kind == SourceUpdate::Changed) {
            clock_.next();
            cur_[0] = in.bits[0];
            return SourceUpdate{SourceUpdate::Changed, cur_, 2};
        }
        if (clock_.next()) {
            cur_[1] = in.bits[1];
            return SourceUpdate{SourceUpdate::Changed, cur_, 2};
        }
        return SourceUpdate{SourceUpdate::Unchanged, cur_, 2};
    }
    size_t q_offset() const override { return q_off_; }
    D& inner() { return data_; }
};
/* type-erased form used by src/bin/modulate.rs:101-107 (`Box<data::Source>` around any inner source) */
class EvenOddOffsetBoxed : public Source {
    std::unique_ptr<Source> data_;
    SymbolClock clock_;
    uint8_t cur_[2] = {0, 0};
    size_t q_off_;
public:
    EvenOddOffsetBoxed(std::unique_ptr<Source> d, size_t samples_per_symbol, size_t bits_per_symbol)
        : data_(std::move(d)), clock_(samples_per_symbol / (bits_per_symbol ? bits_per_symbol : 1)), q_off_(samples_per_symbol / 2)
    {
        if (bits_per_symbol != 2) throw Panic("assertion failed: bits_per_symbol == 2");
        if (samples_per_symbol % bits_per_symbol) throw Panic("assertion failed: samples_per_symbol % bits_per_symbol == 0");
    }
    SourceUpdate next() override
    {
        SourceUpdate in = data_->next();
        if (in.kind == SourceUpdate::Finished) return in;
        if (in.kind == SourceUpdate::Changed) {
            clock_.next();
            cur_[0] = in.bits[0];
            return SourceUpdate{SourceUpdate::Changed, cur_, 2};
        }
        if (clock_.next()) {
            cur_[1] = in.bits[1];
            return SourceUpdate{SourceUpdate::Changed, cur_, 2};
        }
        return SourceUpdate{SourceUpdate::Unchanged, cur_, 2};
    }
    size_t q_offset() const override { return q_off_; }
};
/* data.rs:125-186: ASCII '0'/'1' characters from a stream, whitespace skipped, anything else panics */
class AsciiBits : public Source {
    std::istream& stream_;
    SymbolClock clock_;
    std::vector<uint8_t> bits_;
    bool next_bit(uint8_t* out) /* data.rs:143-162 */
    {
        for (;;) {
            const int c = stream_.get();
            if (c == std::char_traits<char>::eof()) return false;
            /* char::is_whitespace for a byte value (White_Space property, U+0000..U+00FF) */
            if (c == ' ' || (c >= 0x09 && c <= 0x0d) || c == 0x85 || c == 0xa0) continue;
            if (c != '0' && c != '1') throw Panic("assertion failed: (bit as char).is_digit(2)"); /* data.rs:158 */
            *out = (uint8_t)(c - '0');
            return true;
        }
    }
    bool read_bits() /* data.rs:164-174 */
    {
        for (auto& b : bits_)
            if (!next_bit(&b)) return false;
        return true;
    }
public:
    AsciiBits(std::istream& stream, size_t samples_per_symbol, size_t bits_per_symbol)
        : stream_(stream), clock_(samples_per_symbol), bits_(bits_per_symbol, 0) {}
    SourceUpdate next() override /* data.rs:177-189 */
    {
        if (clock_.next()) {
            if (read_bits()) return SourceUpdate{SourceUpdate::Changed, bits_.data(), bits_.size()};
            return SourceUpdate{SourceUpdate::Finished, nullptr, 0};
        }
        return SourceUpdate{SourceUpdate::Unchanged, bits_.data(), bits_.size()};
    }
};
} // namespace data

/* ------------------------------------------------------------------ digital/ */
namespace digital {
struct DigitalPhasor { /* digital/phasor.rs:1-12 */
    virtual ~DigitalPhasor() = default;
    virtual size_t bits_per_symbol() const = 0;
    /* constellation: n_tables() * 2^bps (i,q) pairs; symbol k uses table k % n_tables() (dcqpsk's update()) */
    virtual size_t n_tables() const { return 1; }
    virtual void table(float* out_iq) const = 0;
    /* stateful / time-varying schemes (update() or a sample-dependent i()/q()): evaluated by the phasor
     * kernels from this parameter block instead of a table */
    virtual bool stateful() const { return false; }
    virtual modem_phasor_t phasor_params() const { return modem_phasor_t{}; }
    /* i()/q() of the reference, served from the table (s is ignored by every memoryless scheme) */
    std::pair<float, float> next(size_t /*s*/, const uint8_t* b, size_t symbol_index = 0) const
    {
        const size_t bps = bits_per_symbol(), np = (size_t)1 << bps;
        std::vector<float> t(2 * np * n_tables());
        table(t.data());
        size_t idx = 0;
        for (size_t j = 0; j < bps; ++j) idx = (idx << 1) | (b[j] & 1u); /* digital/util.rs:5-11 */
        const size_t o = 2 * ((symbol_index % n_tables()) * np + idx);
        return {t[o], t[o + 1]};
    }
    float i(size_t s, const uint8_t* b) const { return next(s, b).first; }
    float q(size_t s, const uint8_t* b) const { return next(s, b).second; }
};
inline void ck_table(int rc, const char* what)
{
    if (rc < 0) throw Panic(std::string("assertion failed in ") + what);
}
namespace bask {
struct BASK : DigitalPhasor {
    float amplitude;
    explicit BASK(float a) : amplitude(a) {}
    size_t bits_per_symbol() const override { return 1; }
    void table(float* o) const override { ck_table(modem_const_bask(amplitude, o), "BASK"); }
};
} // namespace bask
namespace bpsk {
struct BPSK : DigitalPhasor {
    float phase, amplitude;
    BPSK(float p, float a) : phase(p), amplitude(a) {}
    size_t bits_per_symbol() const override { return 1; }
    void table(float* o) const override { ck_table(modem_const_bpsk(phase, amplitude, o), "BPSK"); }
};
} // namespace bpsk
namespace qpsk {
struct QPSK : DigitalPhasor {
    float phase, amplitude;
    QPSK(float p, float a) : phase(p), amplitude(a) {}
    size_t bits_per_symbol() const override { return 2; }
    void table(float* o) const override { ck_table(modem_const_qpsk(phase, amplitude, o), "QPSK"); }
};
} // namespace qpsk
namespace qam {
struct QAM : DigitalPhasor {
    size_t bps;
    float phase, amplitude;
    QAM(size_t bits_per_symbol, float p, float a) : bps(bits_per_symbol), phase(p), amplitude(a)
    {
        if (!(bits_per_symbol > 1)) throw Panic("assertion failed: bits_per_symbol > 1"); /* qam.rs:17 */
    }
    size_t bits_per_symbol() const override { return bps; }
    void table(float* o) const override { ck_table(modem_const_qam((uint32_t)bps, phase, amplitude, o), "QAM"); }
};
} // namespace qam
namespace mpsk {
struct MPSK : DigitalPhasor {
    size_t bps;
    float phase_offset, amplitude;
    MPSK(size_t bits_per_symbol, float po, float a) : bps(bits_per_symbol), phase_offset(po), amplitude(a) {}
    size_t bits_per_symbol() const override { return bps; }
    void table(float* o) const override { ck_table(modem_const_mpsk((uint32_t)bps, phase_offset, amplitude, o), "MPSK"); }
};
} // namespace mpsk
namespace oqpsk {
struct OQPSK : DigitalPhasor {
    float amplitude;
    explicit OQPSK(float a) : amplitude(a) {}
    size_t bits_per_symbol() const override { return 2; }
    void table(float* o) const override { ck_table(modem_const_oqpsk(amplitude, o), "OQPSK"); }
};
} // namespace oqpsk
namespace dcqpsk {
struct DCQPSK : DigitalPhasor {
    float amplitude;
    explicit DCQPSK(float a) : amplitude(a) {}
    size_t bits_per_symbol() const override { return 2; }
    size_t n_tables() const override { return 2; }
    void table(float* o) const override { ck_table(modem_const_dcqpsk(amplitude, o), "DCQPSK"); }
};
} // namespace dcqpsk
namespace apsk {
struct Ring : modem_ring_t { /* apsk.rs:60-82 */
    Ring(uint8_t start_, uint8_t end_, float radius_, float phase_)
    {
        if (!(radius_ >= 0.0f && radius_ <= 1.0f)) throw Panic("assertion failed: radius >= 0.0 && radius <= 1.0"); /* apsk.rs:74 */
        start = start_;
        end = end_;
        radius = radius_;
        phase = phase_;
    }
};
struct APSK : DigitalPhasor {
    float amplitude;
    size_t bps;
    std::vector<modem_ring_t> rings;
    APSK(float a, size_t bits_per_symbol, const std::vector<Ring>& r) : amplitude(a), bps(bits_per_symbol), rings(r.begin(), r.end())
    {
        std::vector<float> t(2 * ((size_t)1 << bps));
        if (modem_const_apsk(amplitude, (uint32_t)bps, rings.data(), rings.size(), t.data()) < 0)
            throw Panic("assertion failed: verify(&rings[..], bits_per_symbol)"); /* apsk.rs:26 */
    }
    size_t bits_per_symbol() const override { return bps; }
    void table(float* o) const override { ck_table(modem_const_apsk(amplitude, (uint32_t)bps, rings.data(), rings.size(), o), "APSK"); }
};
} // namespace apsk

/* the memoryless `-m` names of src/bin/modulate.rs:74-95 */
inline std::unique_ptr<DigitalPhasor> by_name(const std::string& dmod)
{
    const float A = 1.0f, PI = 3.14159265358979323846264338327950288f;
    if (dmod == "bask") return std::make_unique<bask::BASK>(A);
    if (dmod == "bpsk") return std::make_unique<bpsk::BPSK>(PI / 4.0f, A);
    if (dmod == "qpsk") return std::make_unique<qpsk::QPSK>(0.0f, A);
    if (dmod == "qam16") return std::make_unique<qam::QAM>(4, 0.0f, A);
    if (dmod == "qam256") return std::make_unique<qam::QAM>(8, 0.0f, A);
    if (dmod == "16psk") return std::make_unique<mpsk::MPSK>(4, 0.0f, A);
    if (dmod == "oqpsk") return std::make_unique<oqpsk::OQPSK>(A);
    if (dmod == "dcqpsk") return std::make_unique<dcqpsk::DCQPSK>(A);
    if (dmod == "16apsk")
        return std::make_unique<apsk::APSK>(A, 4, std::vector<apsk::Ring>{apsk::Ring(0, 4, 0.5f, PI / 4.0f), apsk::Ring(4, 16, 1.0f, PI / 12.0f)});
    throw Panic("invalid digital modulation"); /* modulate.rs:94 */
}
/* ---- the stateful schemes (TX only: the reference has no receiver for them) ---- */
struct StatefulPhasor : DigitalPhasor {
    modem_phasor_t p{};
    StatefulPhasor() { p.struct_size = sizeof p; }
    size_t bits_per_symbol() const override { return p.bits_per_symbol; }
    bool stateful() const override { return true; }
    modem_phasor_t phasor_params() const override { return p; }
    void table(float* o) const override { std::memset(o, 0, sizeof(float) * 2 * ((size_t)1 << p.bits_per_symbol)); }
};
namespace bfsk {
struct BFSK : StatefulPhasor { /* bfsk.rs:14-21 */
    BFSK(const freq::Freq& d, float a)
    {
        p.kind = MODEM_PHASOR_BFSK;
        p.bits_per_symbol = 1;
        p.deviation = d.sample_freq();
        p.amplitude = a;
    }
};
} // namespace bfsk
namespace mfsk {
struct DefaultMap {};  /* mfsk.rs:11-27 */
struct IncreaseMap {}; /* mfsk.rs:29-35 */
struct MFSK : StatefulPhasor { /* mfsk.rs:47-58 */
    MFSK(size_t bits_per_symbol, const freq::Freq& deviation, float amplitude, IncreaseMap) { init(bits_per_symbol, deviation, amplitude, 1); }
    MFSK(size_t bits_per_symbol, const freq::Freq& deviation, float amplitude, DefaultMap) { init(bits_per_symbol, deviation, amplitude, 0); }
private:
    void init(size_t bps, const freq::Freq& d, float a, uint32_t inc)
    {
        p.kind = MODEM_PHASOR_MFSK;
        p.bits_per_symbol = (uint32_t)bps;
        p.deviation = d.sample_freq();
        p.amplitude = a;
        p.mfsk_increase_map = inc;
    }
};
} // namespace mfsk
namespace cpfsk {
struct CPFSK : StatefulPhasor { /* cpfsk.rs:14-23 */
    CPFSK(size_t bits_per_symbol, const rates::Rates& r, float amplitude, size_t deviation)
    {
        p.kind = MODEM_PHASOR_CPFSK;
        p.bits_per_symbol = (uint32_t)bits_per_symbol;
        p.deviation = freq::Freq(deviation * r.baud_rate / 2, r.sample_rate).sample_freq();
        p.amplitude = amplitude;
    }
};
} // namespace cpfsk
namespace msk {
struct MSK : StatefulPhasor { /* msk.rs:12-19 */
    MSK(float amplitude, size_t samples_per_symbol)
    {
        if (samples_per_symbol % 2) throw Panic("assertion failed: samples_per_symbol % 2 == 0"); /* msk.rs:13 */
        p.kind = MODEM_PHASOR_MSK;
        p.bits_per_symbol = 2;
        p.amplitude = amplitude;
    }
};
} // namespace msk
namespace dmpsk {
struct DMPSK : StatefulPhasor { /* dmpsk.rs:16-23 */
    DMPSK(size_t bits_per_symbol, float amplitude, float phase, float shift)
    {
        p.kind = MODEM_PHASOR_DMPSK;
        p.bits_per_symbol = (uint32_t)bits_per_symbol;
        p.amplitude = amplitude;
        p.phase = phase;
        p.shift = shift;
    }
};
} // namespace dmpsk

/* every `-m` name of src/bin/modulate.rs:74-95 */
inline std::unique_ptr<DigitalPhasor> by_name(const std::string& dmod, const rates::Rates& rates)
{
    const float A = 1.0f, PI = 3.14159265358979323846264338327950288f;
    const size_t sr = rates.sample_rate;
    if (dmod == "bfsk") return std::make_unique<bfsk::BFSK>(freq::Freq(200, sr), A);
    if (dmod == "msk") return std::make_unique<msk::MSK>(A, rates.samples_per_symbol);
    if (dmod == "mfsk") return std::make_unique<mfsk::MFSK>(4, freq::Freq(50, sr), A, mfsk::IncreaseMap{});
    if (dmod == "16cpfsk") return std::make_unique<cpfsk::CPFSK>(4, rates, A, 1);
    if (dmod == "dqpsk") return std::make_unique<dmpsk::DMPSK>(2, A, PI / 4.0f, PI / 2.0f);
    if (dmod == "dbpsk") return std::make_unique<dmpsk::DMPSK>(1, A, PI / 4.0f, PI);
    return by_name(dmod);
}
} // namespace digital

/* ------------------------------------------------------------------ gpu: the batched layer L3' */
namespace gpu {
struct PathConfig {
    size_t samples_per_symbol = 8;
    float sample_freq = 0.0f, phase_offset = 0.0f;
    size_t sample0 = 0, q_offset = 0;
    std::vector<float> tx_taps;            /* empty => rectangular hold (the reference's TX) */
    std::vector<float> rx_taps;            /* Demodulator::new's `lp` */
    size_t decision_delay = 0;
    float rx_gain = 2.0f, slicer_gain = 1.0f; /* demodulator.rs:53-54 */
    uint32_t flags = 0;
};
/* RAII owner of one modem_ctx_t */
class Context {
    modem_ctx_t* ctx_ = nullptr;
    std::vector<float> table_;
    PathConfig cfg_;
    size_t bps_ = 0;
public:
    Context(const digital::DigitalPhasor& phasor, PathConfig cfg, int device = 0) : cfg_(std::move(cfg)), bps_(phasor.bits_per_symbol())
    {
        table_.resize(2 * phasor.n_tables() * ((size_t)1 << bps_));
        phasor.table(table_.data());
        if (cfg_.rx_taps.empty()) {
            fir::FIRFilter lp = fir::lowpass();
            cfg_.rx_taps.assign(lp.coefs, lp.coefs + lp.len);
        }
        modem_cfg_t c{};
        c.struct_size = sizeof c;
        c.bits_per_symbol = (uint32_t)bps_;
        c.samples_per_symbol = (uint32_t)cfg_.samples_per_symbol;
        c.n_tables = (uint32_t)phasor.n_tables();
        c.const_iq = table_.data();
        c.q_offset = (uint32_t)cfg_.q_offset;
        c.sample_freq = cfg_.sample_freq;
        c.phase_offset = cfg_.phase_offset;
        c.sample0 = cfg_.sample0;
        c.n_tx_taps = (uint32_t)cfg_.tx_taps.size();
        c.tx_taps = cfg_.tx_taps.empty() ? nullptr : cfg_.tx_taps.data();
        c.n_rx_taps = (uint32_t)cfg_.rx_taps.size();
        c.rx_taps = cfg_.rx_taps.data();
        c.decision_delay = (uint32_t)cfg_.decision_delay;
        c.rx_gain = cfg_.rx_gain;
        c.slicer_gain = cfg_.slicer_gain;
        c.flags = cfg_.flags;
        check(modem_gpu_create(&ctx_, device, &c), nullptr, "modem_gpu_create");
        if (phasor.stateful()) {
            const modem_phasor_t ph = phasor.phasor_params();
            check(modem_gpu_set_phasor(ctx_, &ph), ctx_, "modem_gpu_set_phasor");
        }
    }
    ~Context() { modem_gpu_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    modem_ctx_t* raw() const { return ctx_; }
    size_t bits_per_symbol() const { return bps_; }
    size_t frame_samples(size_t nbits) const { return modem_gpu_frame_samples(ctx_, nbits); }
    size_t decided_symbols(size_t L) const { return modem_gpu_decided_symbols(ctx_, L); }
};

/* bits [F][nbits] -> tx [F][L] (+ baseband iq); host or device pointers */
struct ModulatorBatch {
    Context& ctx;
    explicit ModulatorBatch(Context& c) : ctx(c) {}
    void modulate(const uint8_t* bits, size_t F, size_t nbits, Complex32* tx, Complex32* iq = nullptr)
    {
        check(modem_gpu_modulate(ctx.raw(), bits, F, nbits, tx, iq), ctx.raw(), "modem_gpu_modulate");
    }
};
struct DemodulatorBatch {
    Context& ctx;
    explicit DemodulatorBatch(Context& c) : ctx(c) {}
    void demodulate(const Complex32* rx, size_t F, size_t L, uint8_t* sym, uint8_t* bits, Complex32* soft = nullptr,
                    Complex32* filt = nullptr, float sigma = 0.0f, uint64_t seed = 0, uint64_t frame0 = 0)
    {
        check(modem_gpu_demodulate(ctx.raw(), rx, F, L, sym, bits, soft, filt, sigma, seed, frame0), ctx.raw(), "modem_gpu_demodulate");
    }
};
/* the two binaries' wire-level batch forms (modulate.rs:118-133, demodulate.rs:29-43) */
struct BinBatch {
    Context& ctx;
    explicit BinBatch(Context& c) : ctx(c) {}
    /* out [F][preamble + L] f32: sync tone then data, real part only */
    void modulate_real(const uint8_t* bits, size_t F, size_t nbits, size_t preamble, float preamble_amplitude, float* out)
    {
        check(modem_gpu_modulate_real(ctx.raw(), bits, F, nbits, preamble, preamble_amplitude, out), ctx.raw(), "modem_gpu_modulate_real");
    }
    /* samples [F][L] i16 / f32 real (or complex analytic); filt [F][L - lock] */
    void demodulate_real(const void* samples, uint32_t fmt, size_t F, size_t L, size_t lock, const fir::FIRFilter* hilbert,
                         float* phase_offset, uint8_t* sym, uint8_t* bits, Complex32* soft, Complex32* filt)
    {
        check(modem_gpu_demodulate_real(ctx.raw(), samples, fmt, F, L, lock, hilbert ? hilbert->coefs : nullptr, hilbert ? hilbert->len : 0,
                                        phase_offset, sym, bits, soft, filt),
              ctx.raw(), "modem_gpu_demodulate_real");
    }
};
struct Loopback {
    Context& ctx;
    explicit Loopback(Context& c) : ctx(c) {}
    /* returns {bit errors, bits compared} */
    std::pair<uint64_t, uint64_t> run(const uint8_t* bits, size_t F, size_t nbits, uint8_t* sym, uint8_t* bits_out,
                                      float sigma = 0.0f, uint64_t seed = 0, uint64_t frame0 = 0, Complex32* tx = nullptr)
    {
        uint64_t cnt[2] = {0, 0};
        check(modem_gpu_loopback(ctx.raw(), bits, F, nbits, sigma, seed, frame0, tx, sym, bits_out, cnt), ctx.raw(), "modem_gpu_loopback");
        return {cnt[0], cnt[1]};
    }
    /* extension: packed payloads, 8 bits per byte, first bit = most significant (1/8 of the bytes over PCIe) */
    std::pair<uint64_t, uint64_t> run_packed(const uint8_t* packed, size_t F, size_t nbits, uint8_t* packed_out, float sigma = 0.0f,
                                             uint64_t seed = 0, uint64_t frame0 = 0)
    {
        uint64_t cnt[2] = {0, 0};
        check(modem_gpu_loopback_packed(ctx.raw(), packed, F, nbits, sigma, seed, frame0, packed_out, cnt), ctx.raw(), "modem_gpu_loopback_packed");
        return {cnt[0], cnt[1]};
    }
};
} // namespace gpu

/* ------------------------------------------------------------------ phasor.rs */
namespace phasor {
struct Phasor { /* phasor.rs:1-3 */
    virtual ~Phasor() = default;
    virtual std::optional<std::pair<float, float>> next(size_t s) = 0;
    /* a constant-amplitude tone is what the CUDA path generates; other analog phasors are not on it */
    virtual bool constant(float* amplitude) const { (void)amplitude; return false; }
};
struct Raw : Phasor { /* phasor.rs:5-24 */
    float amplitude;
    explicit Raw(float a) : amplitude(a) {}
    std::optional<std::pair<float, float>> next(size_t) override { return std::make_pair(amplitude, 0.0f); }
    bool constant(float* a) const override { *a = amplitude; return true; }
};
} // namespace phasor

/* ------------------------------------------------------------------ modulator.rs */
namespace modulator {
struct IQSample { /* modulator.rs:22-49 */
    float i, q;            /* pub fields read by modulate.rs:111-112 (--iq) */
    Complex32 modulated;   /* what modulate() returns: computed by the TX kernel */
    Complex32 modulate() const { return modulated; } /* modulator.rs:45-48 */
};

/* Modulator::new(&mut carrier, phasor) -> endless Iterator<Item = IQSample> (modulator.rs:8-20,51-62): the sync
 * tone of modulate.rs:118-126.  Samples are produced by the tone kernel a block at a time; the borrowed
 * Carrier's counter advances with every sample handed out, exactly like the reference's per-sample next(). */
class Modulator {
    carrier::Carrier& carrier_;
    std::unique_ptr<phasor::Phasor> phasor_;
    std::vector<Complex32> buf_;
    size_t pos_ = 0, block_;
    float amplitude_ = 0.0f;
public:
    Modulator(carrier::Carrier& c, std::unique_ptr<phasor::Phasor> p, size_t block = 4096)
        : carrier_(c), phasor_(std::move(p)), block_(block)
    {
        if (!phasor_->constant(&amplitude_)) throw Panic("Modulator: only phasor::Raw runs on the CUDA path");
    }
    std::optional<IQSample> next()
    {
        if (pos_ >= buf_.size()) {
            digital::qpsk::QPSK dummy(0.0f, 1.0f);
            gpu::PathConfig cfg;
            cfg.sample_freq = carrier_.sample_freq;
            cfg.sample0 = carrier_.sample;
            gpu::Context ctx(dummy, cfg);
            buf_.resize(block_);
            check(modem_gpu_preamble(ctx.raw(), 1, block_, amplitude_, buf_.data()), ctx.raw(), "modem_gpu_preamble");
            pos_ = 0;
        }
        carrier_.sample += 1; /* modulator.rs:55 */
        return IQSample{amplitude_, 0.0f, buf_[pos_++]};
    }
    /* `.take(n)`: n samples in one call */
    std::vector<IQSample> take(size_t n)
    {
        std::vector<IQSample> out;
        out.reserve(n);
        const size_t keep = block_;
        if (pos_ >= buf_.size() && n) block_ = n;
        for (size_t i = 0; i < n; ++i) out.push_back(*next());
        block_ = keep;
        return out;
    }
};

/* DigitalModulator::new(&mut carrier, phasor, src) -> Iterator<Item = IQSample> (modulator.rs:64-101) */
class DigitalModulator {
    std::unique_ptr<data::Source> data_;
    carrier::Carrier& carrier_; /* borrowed mutably, like the reference: the counter carries over */
    std::unique_ptr<digital::DigitalPhasor> phasor_;
    size_t sps_;
    gpu::PathConfig extra_;
    std::vector<Complex32> tx_, iq_;
    size_t pos_ = 0;
    bool started_ = false;

    void run()
    {
        /* drain the Source exactly as the per-sample loop would (data.rs semantics stay on the host):
         * every Changed update at a symbol edge contributes bits_per_symbol bits */
        std::vector<uint8_t> bits;
        const size_t bps = phasor_->bits_per_symbol();
        size_t n = 0;
        const size_t q_off = data_->q_offset();
        std::vector<uint8_t> cur(bps, 0);
        for (;; ++n) {
            data::SourceUpdate u = data_->next();
            if (u.kind == data::SourceUpdate::Finished) break;
            if (n % sps_ == 0) { /* symbol edge: the I-rail bits are current, remember the slot */
                bits.insert(bits.end(), u.bits, u.bits + bps);
            } else if (q_off && n % sps_ == q_off) { /* EvenOddOffset: the odd bit arrives half a symbol late */
                bits[bits.size() - bps + 1] = u.bits[1];
            }
        }
        gpu::PathConfig cfg = extra_;
        cfg.samples_per_symbol = sps_;
        cfg.sample_freq = carrier_.sample_freq;
        cfg.sample0 = carrier_.sample;
        cfg.q_offset = q_off;
        gpu::Context ctx(*phasor_, cfg);
        const size_t L = ctx.frame_samples(bits.size());
        tx_.resize(L);
        iq_.resize(L);
        if (L) gpu::ModulatorBatch(ctx).modulate(bits.data(), 1, bits.size(), tx_.data(), iq_.data());
        carrier_.sample += L + 1; /* the reference calls carrier.next() once more before seeing Finished (modulator.rs:86-89) */
        started_ = true;
    }

public:
    DigitalModulator(carrier::Carrier& c, std::unique_ptr<digital::DigitalPhasor> phasor, std::unique_ptr<data::Source> src,
                     size_t samples_per_symbol, gpu::PathConfig extra = {})
        : data_(std::move(src)), carrier_(c), phasor_(std::move(phasor)), sps_(samples_per_symbol), extra_(std::move(extra)) {}
    std::optional<IQSample> next() /* Iterator::next */
    {
        if (!started_) run();
        if (pos_ >= tx_.size()) return std::nullopt;
        IQSample s{iq_[pos_].re, iq_[pos_].im, tx_[pos_]};
        ++pos_;
        return s;
    }
};
} // namespace modulator

/* ------------------------------------------------------------------ demodulator.rs */
namespace demodulator {
constexpr size_t LOCK_SAMPLES = 64; /* demodulator.rs:5 */

/* Demodulator::new(carrier, sig, lp) -> Iterator<Item = (f32, f32)> (demodulator.rs:7-56).
 * S is any callable returning std::optional<Complex32> (the reference's Iterator<Item = Complex<f32>>). */
template <class S>
class Demodulator {
    carrier::Carrier carrier_; /* owned copy (Carrier: Copy, carrier.rs:3) */
    S sig_;
    fir::FIRFilter lp_;
    float phase_offset_ = 0.0f; /* pll.phase_offset */
    std::vector<Complex32> filt_;
    size_t pos_ = 0;
    bool started_ = false;

    void run()
    {
        std::vector<Complex32> rx;
        while (auto x = sig_()) rx.push_back(*x);
        digital::qpsk::QPSK dummy(0.0f, 1.0f); /* the mapper is irrelevant for the raw (I,Q) stream */
        gpu::PathConfig cfg;
        cfg.samples_per_symbol = 1;
        cfg.sample_freq = carrier_.sample_freq;
        cfg.sample0 = carrier_.sample;
        cfg.phase_offset = phase_offset_;
        cfg.rx_taps.assign(lp_.coefs, lp_.coefs + lp_.len);
        cfg.decision_delay = 0;
        gpu::Context ctx(dummy, cfg);
        filt_.resize(rx.size());
        if (!rx.empty()) gpu::DemodulatorBatch(ctx).demodulate(rx.data(), 1, rx.size(), nullptr, nullptr, nullptr, filt_.data());
        carrier_.sample += rx.size();
        started_ = true;
    }

public:
    template <class F>
    Demodulator(carrier::Carrier c, S sig, F lp) : carrier_(c), sig_(std::move(sig)), lp_(lp()) {} /* lp is called like the reference's closure */
    /* demodulator.rs:32-36: LOCK_SAMPLES analytic samples drive the PLL (pll.rs:16-22) on the GPU; the carrier
     * counter moves past them */
    void lock_phase()
    {
        std::vector<Complex32> head;
        for (size_t i = 0; i < LOCK_SAMPLES; ++i) {
            auto x = sig_();
            if (!x) throw Panic("called `Option::unwrap()` on a `None` value"); /* demodulator.rs:34 */
            head.push_back(*x);
        }
        digital::qpsk::QPSK dummy(0.0f, 1.0f);
        gpu::PathConfig cfg;
        cfg.samples_per_symbol = 1;
        cfg.sample_freq = carrier_.sample_freq;
        cfg.sample0 = carrier_.sample;
        cfg.rx_taps.assign(lp_.coefs, lp_.coefs + lp_.len);
        gpu::Context ctx(dummy, cfg);
        check(modem_gpu_lock_phase(ctx.raw(), head.data(), MODEM_SAMPLES_C32, 1, head.size(), nullptr, 0, LOCK_SAMPLES, &phase_offset_),
              ctx.raw(), "modem_gpu_lock_phase");
        carrier_.sample += LOCK_SAMPLES;
    }
    float phase_offset() const { return phase_offset_; } /* pll.phase_offset */
    void set_phase_offset(float po) { phase_offset_ = po; }
    std::optional<std::pair<float, float>> next()
    {
        if (!started_) run();
        if (pos_ >= filt_.size()) return std::nullopt;
        auto v = filt_[pos_++];
        return std::make_pair(v.re, v.im);
    }
};

/* The composition of src/bin/demodulate.rs:29-41 for a REAL input stream: `input.map(|x| Complex::new(x,
 * hfir.add(x)))` feeding Demodulator::new(carrier, analytic, lowpass), then lock_phase().  The Hilbert FIR is a
 * filter description here (the reference evaluates it per sample inside the closure); it runs in the lock
 * kernel, and one library call does lock + demodulation.  R is a callable returning std::optional<T>, T = int16_t
 * (`iter_16`, src/bin/util.rs) or float. */
template <class T>
class RealDemodulator {
    static_assert(std::is_same<T, int16_t>::value || std::is_same<T, float>::value, "i16 or f32 samples");
    carrier::Carrier carrier_;
    std::vector<T> x_;
    fir::FIRFilter hilbert_, lp_;
    bool lock_ = false, started_ = false;
    float phase_offset_ = 0.0f;
    std::vector<Complex32> filt_;
    size_t pos_ = 0;
    void run()
    {
        digital::qpsk::QPSK dummy(0.0f, 1.0f);
        gpu::PathConfig cfg;
        cfg.samples_per_symbol = 1;
        cfg.sample_freq = carrier_.sample_freq;
        cfg.sample0 = carrier_.sample;
        cfg.rx_taps.assign(lp_.coefs, lp_.coefs + lp_.len);
        gpu::Context ctx(dummy, cfg);
        const size_t lock = lock_ ? LOCK_SAMPLES : 0;
        if (x_.size() < lock) throw Panic("called `Option::unwrap()` on a `None` value"); /* demodulator.rs:34 */
        filt_.resize(x_.size() - lock);
        const uint32_t fmt = std::is_same<T, int16_t>::value ? MODEM_SAMPLES_I16 : MODEM_SAMPLES_F32;
        gpu::BinBatch(ctx).demodulate_real(x_.data(), fmt, 1, x_.size(), lock, &hilbert_, lock ? &phase_offset_ : nullptr, nullptr, nullptr,
                                           nullptr, filt_.data());
        carrier_.sample += x_.size();
        started_ = true;
    }
public:
    template <class R, class F>
    RealDemodulator(carrier::Carrier c, R real_sig, fir::FIRFilter hilbert, F lp) : carrier_(c), hilbert_(hilbert), lp_(lp())
    {
        while (auto v = real_sig()) x_.push_back(*v);
    }
    RealDemodulator(carrier::Carrier c, std::vector<T> samples, fir::FIRFilter hilbert, fir::FIRFilter lp)
        : carrier_(c), x_(std::move(samples)), hilbert_(hilbert), lp_(lp) {}
    void lock_phase() { lock_ = true; } /* evaluated with the first next() */
    float phase_offset()
    {
        if (!started_) run();
        return phase_offset_;
    }
    std::optional<std::pair<float, float>> next()
    {
        if (!started_) run();
        if (pos_ >= filt_.size()) return std::nullopt;
        auto v = filt_[pos_++];
        return std::make_pair(v.re, v.im);
    }
};
} // namespace demodulator

/* ------------------------------------------------------------------ src/bin helpers */
namespace bin {
/* Rust's `{}` for f32 (demodulate.rs:42): the shortest decimal string that round-trips, always positional
 * (never an exponent), no trailing ".0", "NaN" / "inf" / "-inf". */
inline std::string display_f32(float v);
/* src/bin/util.rs:3-37 Read16/Iter16: native-endian i16 words; a trailing odd byte ends the stream */
inline std::vector<int16_t> read_all_i16(std::istream& in)
{
    std::vector<char> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::vector<int16_t> out(raw.size() / 2);
    if (!out.empty()) std::memcpy(out.data(), raw.data(), out.size() * 2);
    return out;
}
} // namespace bin

inline std::string bin::display_f32(float v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0.0f) return std::signbit(v) ? "-0" : "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific); /* shortest round-trip digits */
    std::string sci(buf, r.ptr);
    std::string out;
    size_t i = 0;
    if (sci[0] == '-') {
        out = "-";
        i = 1;
    }
    const size_t e = sci.find('e');
    std::string digits;
    for (size_t k = i; k < e; ++k)
        if (sci[k] != '.') digits.push_back(sci[k]);
    const int exp10 = std::atoi(sci.c_str() + e + 1); /* value = d.ddd * 10^exp10 */
    const int nd = (int)digits.size();
    if (exp10 >= nd - 1) {
        out += digits + std::string((size_t)(exp10 - (nd - 1)), '0');
    } else if (exp10 >= 0) {
        out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
    } else {
        out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    }
    return out;
}

} // namespace modem
