// Test driver for the SDR++-shaped blocks (built against the stand-in core headers in oracle/sdrpp_standin):
// wires  source stream -> PI4DQPSK -> DQPSKSymbolExtractor -> BitUnpacker  with every block start()ed on its
// own worker thread like src/main.cpp:105-110 does, feeds a capture in SDR++-sized buffers and writes the
// unpacked bits to a file.  tests/test_host_block.py compares that file with the oracle's bits.
//   usage: test_host_block <in.f32 (interleaved IQ)> <out.bits> <buffer_samples>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
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
    // argv[4] == "rechunk": a thread between the demodulator and the extractor re-cuts the symbol stream into buffers of
    // 1000 symbols (what a dsp::buffer::Reshaper does): the blocks must not depend on buffer boundaries
    const bool rechunk = argc > 4 && !strcmp(argv[4], "rechunk");
    const bool cycle = argc > 4 && !strcmp(argv[4], "cycle");
    dsp::stream<dsp::complex_t> recut;
    extractor.init(rechunk ? &recut : &demod.out);           // the reference's signatures (src/main.cpp:90-91)
    unpacker.init(&extractor.out);
    if (!demod.ok()) { fprintf(stderr, "demodulator not usable: %s\n", demod.lastError()); return 3; }
    if (cycle) {
        // enable() / disable() / enable() (src/main.cpp:105-114,132-163) with a buffer in flight: the second run must deliver
        // a clean stream again (nothing stale, no worker lost)
        demod.start(); extractor.start(); unpacker.start();
        const int n0 = std::min<int>(chunk, (int)iq.size());
        memcpy(vfo.writeBuf, iq.data(), sizeof(dsp::complex_t) * (size_t)n0);
        vfo.swap(n0);                                         // consumed by the demodulator, its results are left in flight
        std::this_thread::sleep_for(std::chrono::milliseconds(300));
        demod.stop(); extractor.stop(); unpacker.stop();
        demod.reset();
    }
    demod.start();
    extractor.start();
    unpacker.start();
    std::thread cutter;
    if (rechunk) {
        cutter = std::thread([&] {
            std::vector<dsp::complex_t> pend;
            while (true) {
                int n = demod.out.read();
                if (n < 0) { break; }
                pend.insert(pend.end(), demod.out.readBuf, demod.out.readBuf + n);
                demod.out.flush();
                while (pend.size() >= 1000) {
                    memcpy(recut.writeBuf, pend.data(), sizeof(dsp::complex_t) * 1000);
                    if (!recut.swap(1000)) { return; }
                    pend.erase(pend.begin(), pend.begin() + 1000);
                }
            }
        });
    }

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
    if (rechunk) { demod.out.stopReader(); recut.stopWriter(); cutter.join(); }
    sink.join();

    FILE* o = fopen(argv[2], "wb");
    fwrite(bits.data(), 1, bits.size(), o);
    fclose(o);
    printf("bits %zu sync %d standarderr %f last_error '%s'\n", bits.size(), (int)extractor.sync, extractor.standarderr, demod.lastError());
    return 0;
}
