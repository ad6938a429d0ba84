/*
 * loopback.cpp -- the `src/bin`-style caller the north star names: a random payload through
 * modem::modulator::DigitalModulator -> modem::demodulator::Demodulator, then
 * the batched form (modem::gpu::Loopback) on many frames.  Same composition a Rust caller of
 * crate `modem` would write; every sample is computed by the CUDA library.
 *
 *   loopback [-m MOD] [-r RATE] [-b RATE] [-c FREQ] [-n BITS] [-f FRAMES]
 * Defaults follow src/bin/modulate.rs:44-58 (sr 10000, baud 220, carrier 1000 Hz).
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <random>
#include <string>

#include "../modem.hpp"

int main(int argc, char** argv)
{
    std::string dmod = "qpsk";
    size_t sr = 10000, br = 220, cf = 1000, nbits = 1 << 20, frames = 64;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string f = argv[i];
        if (f == "-m") dmod = argv[i + 1];
        else if (f == "-r") sr = strtoull(argv[i + 1], nullptr, 10);
        else if (f == "-b") br = strtoull(argv[i + 1], nullptr, 10);
        else if (f == "-c") cf = strtoull(argv[i + 1], nullptr, 10);
        else if (f == "-n") nbits = strtoull(argv[i + 1], nullptr, 10);
        else if (f == "-f") frames = strtoull(argv[i + 1], nullptr, 10);
    }
    try {
        if (!(cf < sr / 2)) throw modem::Panic("assertion failed: cf < sr / 2"); /* modulate.rs:68 */
        const modem::rates::Rates rates(br, sr);
        const size_t sps = rates.samples_per_symbol;
        auto phasor = modem::digital::by_name(dmod);
        const size_t bps = phasor->bits_per_symbol();
        std::mt19937_64 rng(0x5EED0001);
        std::vector<uint8_t> bits(nbits);
        for (auto& b : bits) b = (uint8_t)(rng() & 1);

        /* ---- streaming API, exactly as the reference's bins compose it.  One Carrier counts samples
         * in a usize but evaluates `s as f32` (carrier.rs:18): past 2^24 samples the counter is
         * quantised and the reference's own round trip starts making bit errors (the oracle shows
         * the same, tests/test_gpu_parity.py::test_parity_across_2p24).  A single stream is therefore
         * kept below 2^24 samples; the full payload goes through the batched leg as frames. */
        const size_t stream_bits = std::min(nbits, (((size_t)1 << 24) / sps - 64) * bps);
        modem::carrier::Carrier carrier_tx(modem::freq::Freq(cf, sr));
        modem::modulator::DigitalModulator mod(carrier_tx, modem::digital::by_name(dmod),
                                               std::make_unique<modem::data::Bits>(bits.data(), stream_bits, sps, bps), sps);
        auto sig = [&]() -> std::optional<modem::Complex32> {
            auto s = mod.next();
            if (!s) return std::nullopt;
            return s->modulate();
        };
        modem::demodulator::Demodulator<decltype(sig)> demod(modem::carrier::Carrier(modem::freq::Freq(cf, sr)), sig, modem::fir::lowpass);
        const size_t delay = 31 + sps / 2;
        size_t n = 0, k = 0, errors = 0, decided = 0;
        const auto t0 = std::chrono::steady_clock::now();
        while (auto iq = demod.next()) {
            if (n >= delay && (n - delay) % sps == 0 && bps == 2) { /* QPSK slicer: b0 = I > 0, b1 = Q > 0 */
                errors += (uint8_t)(iq->first > 0) != bits[2 * k];
                errors += (uint8_t)(iq->second > 0) != bits[2 * k + 1];
                decided += 2;
                ++k;
            }
            ++n;
        }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("streaming API: %zu samples, %zu bits decided, %zu bit errors, %.1f Msamples/s (one frame, host buffers)\n", n, decided,
               errors, n / dt / 1e6);
        if (bps == 2 && errors) return 1;

        /* ---- batched API: many frames, slicer + error count on the device */
        modem::gpu::PathConfig cfg;
        cfg.samples_per_symbol = sps;
        cfg.sample_freq = modem::freq::Freq(cf, sr).sample_freq();
        cfg.decision_delay = delay;
        {
            modem::fir::FIRFilter lp = modem::fir::lowpass();
            float g = 0.0f;
            for (size_t i = 0; i < lp.len; ++i) g += lp.coefs[i];
            cfg.slicer_gain = g;
        }
        modem::gpu::Context ctx(*phasor, cfg);
        const size_t fb = (nbits / frames / bps) * bps;
        const size_t L = ctx.frame_samples(fb), K = ctx.decided_symbols(L);
        std::vector<uint8_t> out(frames * K * bps);
        const auto t1 = std::chrono::steady_clock::now();
        auto cnt = modem::gpu::Loopback(ctx).run(bits.data(), frames, fb, nullptr, out.data());
        const double dt2 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
        printf("batched API:   %zu frames x %zu samples, %llu bits compared, %llu bit errors, %.1f Msamples/s (host buffers)\n", frames, L,
               (unsigned long long)cnt.second, (unsigned long long)cnt.first, frames * L / dt2 / 1e6);
        return cnt.first == 0 ? 0 : 1;
    } catch (const modem::Panic& e) {
        fprintf(stderr, "panicked: %s\n", e.what());
        return 101; /* Rust's panic exit code */
    }
}
