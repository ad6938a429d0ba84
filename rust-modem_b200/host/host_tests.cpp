/*
 * host_tests.cpp -- the reference's unit tests restated on the C++ mirror (modem.hpp), plus
 * GPU round trips through the mirrored streaming API.
 *   host_tests --cpu   source / clock / mapper tests (no device needed)
 *   host_tests --gpu   DigitalModulator -> Demodulator through the CUDA library
 */
#include <cstdio>
#include <cstring>
#include <random>
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
    if (gpu) {
        try {
            test_gpu_roundtrip("qpsk", 1250, 2500);
            test_gpu_roundtrip("qpsk", 220, 1000); /* the reference's default rates: sps 45 */
            test_gpu_roundtrip("oqpsk", 1250, 2500);
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
