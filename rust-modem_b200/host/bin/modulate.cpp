/*
 * modulate.cpp -- mirror of /root/reference/src/bin/modulate.rs on the CUDA path: "Modulate the bits on stdin
 * to a waveform on stdout".  Same options, defaults, assertions and wire formats:
 *   stdin   ASCII '0'/'1' characters, whitespace ignored (data.rs:125-186)
 *   stdout  f32 little-endian: real part of the modulated waveform (modulate.rs:128-133), preceded by
 *           `sr / cf * CYCLES - 1` samples of carrier sync tone with -p (modulate.rs:118-126); with --iq the
 *           interleaved baseband (i, q) pairs (modulate.rs:109-116)
 * The composition below follows the reference line by line; the per-sample work behind the iterators is done
 * by the CUDA library in one call per stream.
 */
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "../modem.hpp"

using namespace modem;

static const float AMPLITUDE = 1.0f; /* modulate.rs:14 */

static void usage()
{
    std::printf("Usage: modulate [options]\n\n    Modulate the bits on stdin to a waveform on stdout\n\nOptions:\n"
                "    -h, --help          show usage\n    -m MOD              digital modulation to use\n"
                "    -r RATE             sample rate (samples/sec)\n    -b RATE             baud rate (symbols/sec)\n"
                "    -c FREQ             carrier frequency (Hz)\n    -p CYCLES           preamble cycles\n"
                "        --iq            output raw IQ samples\n");
}

static size_t parse(const char* s, const char* what)
{
    char* end = nullptr;
    const unsigned long long v = std::strtoull(s, &end, 10);
    if (!*s || *end) throw Panic(what);
    return (size_t)v;
}

static void write_f32(float v) { std::fwrite(&v, sizeof v, 1, stdout); } /* LittleEndian on every supported host */

int main(int argc, char** argv)
{
    try {
        std::string dmod;
        bool have_m = false, iq = false;
        size_t sr = 10000, br = 220, cf = 1000, pc = 0; /* modulate.rs:44-66 */
        bool have_p = false;
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            auto val = [&](const char* what) -> std::string {
                if (a.size() > 2) return a.substr(2);
                if (i + 1 >= argc) throw Panic(what);
                return argv[++i];
            };
            if (a == "-h" || a == "--help") {
                usage();
                return 0;
            } else if (a == "--iq") {
                iq = true;
            } else if (a.rfind("-m", 0) == 0) {
                dmod = val("digital modulation is required");
                have_m = true;
            } else if (a.rfind("-r", 0) == 0) {
                sr = parse(val("invalid sample rate").c_str(), "invalid sample rate");
            } else if (a.rfind("-b", 0) == 0) {
                br = parse(val("invalid baud rate").c_str(), "invalid baud rate");
            } else if (a.rfind("-c", 0) == 0) {
                cf = parse(val("invalid carrier frequency").c_str(), "invalid carrier frequency");
            } else if (a.rfind("-p", 0) == 0) {
                pc = parse(val("invalid preamble cycles").c_str(), "invalid preamble cycles");
                have_p = true;
            } else {
                throw Panic("unrecognized option"); /* getopts parse().unwrap() */
            }
        }
        if (!have_m) throw Panic("digital modulation is required"); /* modulate.rs:41 */
        if (have_p && sr % cf != 0) throw Panic("assertion failed: sr % cf == 0"); /* modulate.rs:62 */
        if (!(cf < sr / 2)) throw Panic("assertion failed: cf < sr / 2");           /* modulate.rs:68 */

        const rates::Rates rates(br, sr);
        carrier::Carrier carrier{freq::Freq(cf, sr)};
        auto phasor = digital::by_name(dmod, rates); /* modulate.rs:74-95 */
        const size_t bps = phasor->bits_per_symbol();

        auto bits = std::make_unique<data::AsciiBits>(std::cin, rates.samples_per_symbol, bps); /* modulate.rs:98-99 */
        std::unique_ptr<data::Source> src;
        if (dmod == "msk" || dmod == "oqpsk") /* modulate.rs:101-107 */
            src = std::make_unique<data::EvenOddOffsetBoxed>(std::move(bits), rates.samples_per_symbol, bps);
        else
            src = std::move(bits);

        if (iq) { /* modulate.rs:109-116 */
            modulator::DigitalModulator m(carrier, std::move(phasor), std::move(src), rates.samples_per_symbol);
            while (auto s = m.next()) {
                write_f32(s->i);
                write_f32(s->q);
            }
            return 0;
        }
        if (pc > 0) { /* modulate.rs:118-126 */
            modulator::Modulator preamble(carrier, std::make_unique<phasor::Raw>(AMPLITUDE));
            for (const auto& s : preamble.take(sr / cf * pc - 1)) write_f32(s.modulate().re);
        }
        modulator::DigitalModulator digi(carrier, std::move(phasor), std::move(src), rates.samples_per_symbol); /* :128-133 */
        while (auto s = digi.next()) write_f32(s->modulate().re);
        return 0;
    } catch (const Panic& p) {
        std::fprintf(stderr, "thread 'main' panicked at '%s'\n", p.what());
        return 101;
    }
}
