/*
 * demodulate.cpp -- mirror of /root/reference/src/bin/demodulate.rs on the CUDA path: "Demodulate a waveform on
 * stdin to i/q samples on stdout".
 *   stdin   native-endian i16 samples (src/bin/util.rs:3-37), `x as f32` (demodulate.rs:29)
 *   stdout  one line `i:{}\tq:{}` per demodulated sample (demodulate.rs:41-43), Rust float formatting
 * Sample rate 10000 and carrier 900 Hz are the reference's constants (demodulate.rs:9,36); -b is accepted and,
 * as in the reference, unused.  Hilbert FIR + 64-sample PLL lock + the two low-pass FIRs run in the CUDA library.
 */
#include <cstdio>
#include <iostream>
#include <string>

#include "../modem.hpp"

using namespace modem;

static const size_t SAMPLE_RATE = 10000; /* demodulate.rs:9 */

int main(int argc, char** argv)
{
    try {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            if (a == "-h" || a == "--help") {
                std::printf("Usage: demodulate [options]\n\n    Demodulate a waveform on stdin to i/q samples on stdout\n\nOptions:\n"
                            "    -h, --help          show usage\n    -b RATE             baud rate (symbols/sec)\n");
                return 0;
            } else if (a.rfind("-b", 0) == 0) {
                if (a.size() == 2 && ++i >= argc) throw Panic("Argument to option 'b' missing");
            } else {
                throw Panic("unrecognized option");
            }
        }
        std::vector<int16_t> input = bin::read_all_i16(std::cin); /* demodulate.rs:29 */
        const freq::Freq carrier_freq(900, SAMPLE_RATE);            /* demodulate.rs:36 */
        demodulator::RealDemodulator<int16_t> demod(carrier::Carrier(carrier_freq), std::move(input), fir::hilbert(), fir::lowpass());
        demod.lock_phase(); /* demodulate.rs:39 */
        std::string line;
        while (auto v = demod.next()) { /* demodulate.rs:41-43 */
            line = "i:" + bin::display_f32(v->first) + "\tq:" + bin::display_f32(v->second) + "\n";
            std::fwrite(line.data(), 1, line.size(), stdout);
        }
        return 0;
    } catch (const Panic& p) {
        std::fprintf(stderr, "thread 'main' panicked at '%s'\n", p.what());
        return 101;
    }
}
