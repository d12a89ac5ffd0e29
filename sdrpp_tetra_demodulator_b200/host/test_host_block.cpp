// Test driver for the SDR++-shaped blocks (built against the stand-in core headers in oracle/sdrpp_standin):
// wires  source stream -> PI4DQPSK -> DQPSKSymbolExtractor -> BitUnpacker  with every block start()ed on its
// own worker thread like src/main.cpp:105-110 does, feeds a capture in SDR++-sized buffers and writes the
// unpacked bits to a file.  tests/test_host_block.py compares that file with the oracle's bits.
//   usage: test_host_block <in.f32 (interleaved IQ)> <out.bits> <buffer_samples>
#include <stdio.h>
#include <atomic>
#include <thread>
#include <vector>
#include "pi4dqpsk_b200.h"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { return 2; }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<dsp::complex_t> iq((size_t)bytes / sizeof(dsp::complex_t));
    if (fread(iq.data(), sizeof(dsp::complex_t), iq.size(), f) != iq.size()) { return 2; }
    fclose(f);
    const int chunk = atoi(argv[3]);

    // src/main.cpp:78-84
    tdm_config c;
    tdm_default_config(&c);
    dsp::stream<dsp::complex_t> vfo;
    dsp::b200::PI4DQPSK demod;
    dsp::b200::DQPSKSymbolExtractor extractor;
    dsp::b200::BitUnpacker unpacker;
    demod.init(&vfo, c.symbolrate, c.samplerate, c.rrc_tap_count, c.rrc_beta, c.agc_rate, c.costas_bandwidth,
               c.fll_bandwidth, c.omega_gain, c.mu_gain, c.omega_rel_limit);
    extractor.init(&demod.out, &demod);
    unpacker.init(&extractor.out, &extractor);
    demod.start();
    extractor.start();
    unpacker.start();

    std::vector<uint8_t> bits;
    std::atomic<bool> done{ false };
    std::thread sink([&] {
        while (true) {
            int n = unpacker.out.read();
            if (n < 0) { break; }
            bits.insert(bits.end(), unpacker.out.readBuf, unpacker.out.readBuf + n);
            unpacker.out.flush();
        }
        done = true;
    });

    size_t pos = 0;
    while (pos < iq.size()) {
        int n = (int)std::min<size_t>((size_t)chunk, iq.size() - pos);
        memcpy(vfo.writeBuf, &iq[pos], sizeof(dsp::complex_t) * (size_t)n);
        if (!vfo.swap(n)) { break; }
        pos += (size_t)n;
    }
    // let the pipeline drain: every buffer handed to swap() has been read when the next swap() returns
    std::this_thread::sleep_for(std::chrono::milliseconds(500));
    demod.stop();
    extractor.stop();
    unpacker.stop();
    unpacker.out.stopReader();
    sink.join();

    FILE* o = fopen(argv[2], "wb");
    fwrite(bits.data(), 1, bits.size(), o);
    fclose(o);
    printf("bits %zu sync %d standarderr %f last_error '%s'\n", bits.size(), (int)extractor.sync, extractor.standarderr, demod.lastError());
    return 0;
}
