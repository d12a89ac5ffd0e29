// Test driver for dsp::b200::BurstSync (built against the stand-in core headers in oracle/sdrpp_standin):
// a source stream of bits (one per byte) -> BurstSync start()ed on its own worker thread, fed in SDR++-sized
// buffers; the burst records it puts on `out` are written to a file.  tests/test_host_block.py compares them
// with the checker's.   usage: test_host_bsync <in.bits> <out.bursts> <buffer_bits>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>
#include "burst_sync_b200.h"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { return 2; }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> bits((size_t)bytes);
    if (fread(bits.data(), 1, bits.size(), f) != bits.size()) { return 2; }
    fclose(f);
    const int chunk = atoi(argv[3]);

    dsp::stream<uint8_t> src;
    dsp::b200::BurstSync bsync;
    bsync.init(&src, 0, true);
    bsync.start();

    std::vector<uint8_t> records;
    std::thread sink([&] {
        while (true) {
            int n = bsync.out.read();
            if (n < 0) { break; }
            records.insert(records.end(), bsync.out.readBuf, bsync.out.readBuf + n);
            bsync.out.flush();
        }
    });
    size_t pos = 0;
    while (pos < bits.size()) {
        int n = (int)std::min<size_t>((size_t)chunk, bits.size() - pos);
        memcpy(src.writeBuf, &bits[pos], (size_t)n);
        if (!src.swap(n)) { break; }
        pos += (size_t)n;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(500));
    bsync.stop();
    bsync.out.stopReader();
    sink.join();

    FILE* o = fopen(argv[2], "wb");
    fwrite(records.data(), 1, records.size(), o);
    fclose(o);
    printf("bursts %zu rx_state %d ts_found %d delivered %llu last_error '%s'\n", records.size() / sizeof(tdm_burst), bsync.rxState(),
           (int)bsync.tsFound(), bsync.burstsDelivered(), bsync.lastError());
    return 0;
}
