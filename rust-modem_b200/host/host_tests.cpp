/*
 * host_tests.cpp -- the reference's unit tests restated on the C++ mirror (modem.hpp), plus
 * GPU round trips through the mirrored streaming API.
 *   host_tests --cpu   source / clock / mapper tests (no device needed)
 *   host_tests --gpu   DigitalModulator -> Demodulator through the CUDA library
 */
#include <cstdio>
#include <cstring>
#include <random>
#include <sstream>
#include <vector>

#include "modem.hpp"

using namespace modem;
static int failures = 0;
#define EXPECT(c)                                                        \
    do {                                                                 \
        if (!(c)) {                                                      \
            printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c);        \
            ++failures;                                                  \
        }                                                                \
    } while (0)

static data::SourceUpdate U(data::SourceUpdate::Kind k, std::vector<uint8_t> b)
{
    static uint8_t store[64][8];
    static int slot = 0;
    uint8_t* p = store[slot++ % 64];
    size_t n = 0;
    for (uint8_t v : b) p[n++] = v;
    return data::SourceUpdate{k, p, n};
}

static void test_symbol_clock() /* data.rs:195-209 */
{
    data::SymbolClock bc(5);
    const bool want[11] = {true, false, false, false, false, true, false, false, false, false, true};
    for (bool w : want) EXPECT(bc.next() == w);
}
static void test_bits() /* data.rs:212-224 */
{
    const uint8_t BITS[] = {1, 0, 1, 1};
    data::Bits ds(BITS, 4, 3, 2);
    using K = data::SourceUpdate;
    EXPECT(ds.next() == U(K::Changed, {1, 0}));
    EXPECT(ds.next() == U(K::Unchanged, {1, 0}));
    EXPECT(ds.next() == U(K::Unchanged, {1, 0}));
    EXPECT(ds.next() == U(K::Changed, {1, 1}));
    EXPECT(ds.next() == U(K::Unchanged, {1, 1}));
    EXPECT(ds.next() == U(K::Unchanged, {1, 1}));
    EXPECT(ds.next().kind == K::Finished);
}
static void test_evenodd() /* data.rs:227-246 */
{
    const uint8_t BITS[] = {1, 1, 1, 0, 0, 1};
    data::EvenOddOffset<data::Bits> eo(data::Bits(BITS, 6, 4, 2), 4, 2);
    using K = data::SourceUpdate;
    const std::vector<std::pair<K::Kind, std::vector<uint8_t>>> want = {
        {K::Changed, {1, 0}}, {K::Unchanged, {1, 0}}, {K::Changed, {1, 1}}, {K::Unchanged, {1, 1}},
        {K::Changed, {1, 1}}, {K::Unchanged, {1, 1}}, {K::Changed, {1, 0}}, {K::Unchanged, {1, 0}},
        {K::Changed, {0, 0}}, {K::Unchanged, {0, 0}}, {K::Changed, {0, 1}}, {K::Unchanged, {0, 1}}};
    for (auto& w : want) EXPECT(eo.next() == U(w.first, w.second));
    EXPECT(eo.next().kind == K::Finished);
}
static void test_mpsk() /* digital/mpsk.rs:50-63 */
{
    digital::mpsk::MPSK mpsk(2, 0.0f, 1.0f);
    const uint8_t b00[] = {0, 0}, b01[] = {0, 1}, b10[] = {1, 0}, b11[] = {1, 1};
    EXPECT(mpsk.i(0, b00) == 1.0f && mpsk.q(0, b00) == 0.0f);
    EXPECT(std::fabs(mpsk.i(0, b01)) < 0.001f && mpsk.q(0, b01) == 1.0f);
    EXPECT(mpsk.i(0, b10) == -1.0f && std::fabs(mpsk.q(0, b10)) < 0.001f);
    EXPECT(std::fabs(mpsk.i(0, b11)) < 0.001f && mpsk.q(0, b11) == -1.0f);
}
static void test_qam() /* digital/qam.rs:69-84 */
{
    digital::qam::QAM qam(4, 0.0f, 6.0f);
    const uint8_t a[] = {0, 0, 0, 0}, b[] = {0, 0, 0, 1}, c[] = {1, 0, 1, 1}, d[] = {1, 1, 1, 1};
    EXPECT(qam.i(0, a) == -3.0f && qam.q(0, a) == -3.0f);
    EXPECT(qam.i(0, b) == -3.0f && qam.q(0, b) == -1.0f);
    EXPECT(qam.i(0, c) == 1.0f && qam.q(0, c) == 3.0f);
    EXPECT(qam.i(0, d) == 3.0f && qam.q(0, d) == 3.0f);
}
static void test_panics()
{
    bool threw = false;
    try { digital::by_name("nope"); } catch (const Panic&) { threw = true; } /* modulate.rs:94 */
    EXPECT(threw);
    threw = false;
    try { digital::qam::QAM q(1, 0.0f, 1.0f); } catch (const Panic&) { threw = true; } /* qam.rs:17 */
    EXPECT(threw);
    threw = false;
    try { digital::apsk::APSK a(1.0f, 4, {digital::apsk::Ring(0, 4, 0.5f, 0.0f)}); } catch (const Panic&) { threw = true; } /* apsk.rs:26 */
    EXPECT(threw);
    EXPECT(rates::Rates(220, 10000).samples_per_symbol == 45); /* rates.rs:16 */
}

static void test_ascii_bits() /* data.rs:125-189 (no reference test: behaviour restated from the lines) */
{
    std::istringstream in("1 0\n1\t1 0");
    data::AsciiBits ab(in, 2, 2);
    using K = data::SourceUpdate;
    EXPECT(ab.next() == U(K::Changed, {1, 0}));
    EXPECT(ab.next() == U(K::Unchanged, {1, 0}));
    EXPECT(ab.next() == U(K::Changed, {1, 1}));
    EXPECT(ab.next() == U(K::Unchanged, {1, 1}));
    EXPECT(ab.next().kind == K::Finished); /* a partial trailing symbol ends the stream (data.rs:164-174) */
    std::istringstream bad("102");
    data::AsciiBits ab2(bad, 1, 1);
    EXPECT(ab2.next() == U(K::Changed, {1}));
    EXPECT(ab2.next() == U(K::Changed, {0}));
    bool threw = false;
    try { ab2.next(); } catch (const Panic&) { threw = true; } /* data.rs:158 assert!(is_digit(2)) */
    EXPECT(threw);
}
static void test_display_f32() /* Rust `{}` for f32, demodulate.rs:42 */
{
    EXPECT(bin::display_f32(1.0f) == "1");
    EXPECT(bin::display_f32(-2.5f) == "-2.5");
    EXPECT(bin::display_f32(0.1f) == "0.1");
    EXPECT(bin::display_f32(0.0f) == "0");
    EXPECT(bin::display_f32(1e-7f) == "0.0000001");
    EXPECT(bin::display_f32(16777216.0f) == "16777216");
    EXPECT(bin::display_f32(1e30f) == "1000000000000000000000000000000");
    EXPECT(bin::display_f32(3.4028235e38f) == "340282350000000000000000000000000000000");
    EXPECT(bin::display_f32(0.99864417f) == "0.9986442");
    EXPECT(bin::display_f32(-7999.1234f) == "-7999.1235");
    EXPECT(bin::display_f32(1.17549435e-38f) == "0.000000000000000000000000000000000000011754944");
    EXPECT(bin::display_f32(std::nanf("")) == "NaN");
    EXPECT(bin::display_f32(-INFINITY) == "-inf");
}
static void test_stateful_by_name() /* modulate.rs:74-95 */
{
    const rates::Rates r(1250, 10000);
    const char* names[] = {"bfsk", "mfsk", "16cpfsk", "msk", "dqpsk", "dbpsk"};
    const size_t bps[] = {1, 4, 4, 2, 2, 1};
    for (int i = 0; i < 6; ++i) {
        auto ph = digital::by_name(names[i], r);
        EXPECT(ph->stateful() && ph->bits_per_symbol() == bps[i]);
        modem_phasor_t ref;
        uint32_t eo = 0;
        EXPECT(modem_phasor_by_name(names[i], 1250, 10000, &ref, &eo) == (int)bps[i]);
        const modem_phasor_t got = ph->phasor_params();
        EXPECT(!std::memcmp(&got, &ref, sizeof ref));
    }
    EXPECT(!digital::by_name("qpsk", r)->stateful());
    bool threw = false;
    try { digital::msk::MSK(1.0f, 45); } catch (const Panic&) { threw = true; } /* msk.rs:13 */
    EXPECT(threw);
}

/* sync tone + QPSK on one Carrier -> i16 wire -> Hilbert/PLL lock -> (I,Q): the two binaries' composition */
static void test_gpu_bin_chain()
{
    const size_t sr = 10000, br = 1250, cf = 1000, pc = 20;
    const rates::Rates r(br, sr);
    std::string ascii;
    std::mt19937 rng(11);
    std::vector<uint8_t> bits(2 * 200);
    for (auto& b : bits) {
        b = rng() & 1;
        ascii += (char)('0' + b);
        ascii += ' ';
    }
    std::istringstream in(ascii);
    carrier::Carrier carrier(freq::Freq(cf, sr));
    std::vector<float> wave;
    modulator::Modulator preamble(carrier, std::make_unique<phasor::Raw>(1.0f));
    const size_t P = sr / cf * pc - 1;
    for (const auto& s : preamble.take(P)) wave.push_back(s.modulate().re);
    EXPECT(carrier.sample == P);
    modulator::DigitalModulator digi(carrier, digital::by_name("qpsk", r), std::make_unique<data::AsciiBits>(in, r.samples_per_symbol, 2),
                                     r.samples_per_symbol);
    while (auto s = digi.next()) wave.push_back(s->modulate().re);
    EXPECT(wave.size() == P + 200 * r.samples_per_symbol);
    std::vector<int16_t> wire;
    for (float v : wave) wire.push_back((int16_t)std::lrintf(v * 8000.0f));
    demodulator::RealDemodulator<int16_t> demod(carrier::Carrier(freq::Freq(cf, sr)), wire, fir::hilbert(), fir::lowpass());
    demod.lock_phase();
    std::vector<std::pair<float, float>> iq;
    while (auto v = demod.next()) iq.push_back(*v);
    EXPECT(iq.size() == wire.size() - demodulator::LOCK_SAMPLES);
    EXPECT(std::fabs(demod.phase_offset()) < 0.6f);
    const size_t sps = r.samples_per_symbol, delay = (P - 64) + 31 + sps / 2;
    size_t errors = 0, decided = 0;
    for (size_t k = 0; k * sps + delay < iq.size(); ++k) {
        errors += (uint8_t)(iq[k * sps + delay].first > 0) != bits[2 * k];
        errors += (uint8_t)(iq[k * sps + delay].second > 0) != bits[2 * k + 1];
        decided += 2;
    }
    printf("  bin chain: tone %zu + %zu samples, lock offset %g, %zu bits decided, %zu errors\n", P, wave.size() - P,
           demod.phase_offset(), decided, errors);
    EXPECT(errors == 0 && decided > 380);
    /* stateful mappers through the streaming API: constant envelope */
    for (const char* name : {"bfsk", "mfsk", "16cpfsk", "msk", "dqpsk"}) {
        carrier::Carrier c2(freq::Freq(cf, sr));
        auto ph = digital::by_name(name, r);
        const size_t bps = ph->bits_per_symbol();
        std::vector<uint8_t> b2(bps * 64);
        for (auto& b : b2) b = rng() & 1;
        std::unique_ptr<data::Source> src = std::make_unique<data::Bits>(b2.data(), b2.size(), sps, bps);
        if (std::string(name) == "msk") src = std::make_unique<data::EvenOddOffsetBoxed>(std::move(src), sps, bps);
        modulator::DigitalModulator m(c2, std::move(ph), std::move(src), sps);
        size_t n = 0;
        float worst = 0.0f;
        while (auto s = m.next()) {
            worst = std::max(worst, std::fabs(std::hypot(s->i, s->q) - 1.0f));
            ++n;
        }
        EXPECT(n == 64 * sps && worst < 1e-6f);
    }
}

static void test_gpu_roundtrip(const char* dmod, size_t sps_br, size_t cf)
{
    const size_t sr = 10000;
    const rates::Rates r(sps_br, sr);
    auto ph = digital::by_name(dmod);
    const size_t bps = ph->bits_per_symbol(), sps = r.samples_per_symbol;
    std::mt19937 rng(7);
    std::vector<uint8_t> bits(bps * 300);
    for (auto& b : bits) b = rng() & 1;
    carrier::Carrier ctx_c(freq::Freq(cf, sr));
    const bool eo = std::string(dmod) == "oqpsk"; /* modulate.rs:101-107 */
    std::unique_ptr<data::Source> src;
    if (eo) src = std::make_unique<data::EvenOddOffset<data::Bits>>(data::Bits(bits.data(), bits.size(), sps, bps), sps, bps);
    else src = std::make_unique<data::Bits>(bits.data(), bits.size(), sps, bps);
    modulator::DigitalModulator mod(ctx_c, digital::by_name(dmod), std::move(src), sps);
    size_t produced = 0;
    auto sig = [&]() -> std::optional<Complex32> {
        auto s = mod.next();
        if (!s) return std::nullopt;
        ++produced;
        return s->modulate();
    };
    demodulator::Demodulator<decltype(sig)> demod(carrier::Carrier(freq::Freq(cf, sr)), sig, fir::lowpass);
    const size_t delay = 31 + sps / 2, qoff = eo ? sps / 2 : 0;
    std::vector<std::pair<float, float>> iq;
    while (auto v = demod.next()) iq.push_back(*v);
    EXPECT(produced == 300 * sps && iq.size() == produced);
    EXPECT(ctx_c.sample == produced + 1); /* the borrowed carrier advanced like the reference's */
    if (bps == 2) {
        size_t errors = 0, decided = 0;
        for (size_t k = 0; k * sps + delay + qoff < iq.size(); ++k) {
            errors += (uint8_t)(iq[k * sps + delay].first > 0) != bits[2 * k];
            errors += (uint8_t)(iq[k * sps + delay + qoff].second > 0) != bits[2 * k + 1];
            decided += 2;
        }
        printf("  %-6s sps %2zu: %zu samples, %zu bits decided, %zu errors\n", dmod, sps, produced, decided, errors);
        EXPECT(errors == 0 && decided > 500);
    }
}

int main(int argc, char** argv)
{
    const bool gpu = argc > 1 && !strcmp(argv[1], "--gpu");
    test_symbol_clock();
    test_bits();
    test_evenodd();
    test_mpsk();
    test_qam();
    test_panics();
    test_ascii_bits();
    test_display_f32();
    test_stateful_by_name();
    if (gpu) {
        try {
            test_gpu_roundtrip("qpsk", 1250, 2500);
            test_gpu_roundtrip("qpsk", 220, 1000); /* the reference's default rates: sps 45 */
            test_gpu_roundtrip("oqpsk", 1250, 2500);
            test_gpu_bin_chain();
        } catch (const Panic& e) {
            printf("FAIL panicked: %s\n", e.what());
            ++failures;
        }
    } else {
        /* no device: the first sample request must fail loudly, never fall back to host arithmetic */
        int n = 0;
        if (modem_gpu_device_count(&n) != MODEM_OK || n == 0) {
            bool threw = false;
            try { test_gpu_roundtrip("qpsk", 1250, 2500); } catch (const Panic&) { threw = true; }
            EXPECT(threw);
        }
    }
    printf("%s (%d failures)\n", failures ? "FAILED" : "ok", failures);
    return failures ? 1 : 0;
}
