// burst_sync_b200.h -- SDR++ block in front of the protocol decoder: the burst synchroniser on the GPU.
//
// In decoder mode the reference wires bitsUnpacker.out into dsp::osmotetradec (src/main.cpp:95), whose process()
// begins with  tetra_burst_sync_in(trs, (uint8_t*)in, count)  (src/dsp/osmotetra_dec.h:183); everything after that
// call (lower MAC, codec, audio pacing) is out of this repo's scope.  dsp::b200::BurstSync is that first step as
// a block of its own: Processor<uint8_t, uint8_t>, input = one bit per byte, output = the bursts the reference
// would have handed to tetra_burst_rx_cb, as packed tdm_burst records (96 bytes each, include/tdm_burst_b200.h).
// A consumer that keeps the reference's lower MAC calls
//     tdm_burst_unpack(rec, bits510); tetra_burst_rx_cb(bits510, 510, (enum tetra_train_seq)rec->train_seq, tms);
// for every record.
//
// One input buffer = one tetra_burst_sync_in call, like the reference -- as long as it holds at most one slot
// (510 bits; SDR++ hands this block a few hundred bits at a time at 36 kS/s).  Longer buffers are fed as
// consecutive 510-bit calls: the reference itself is undefined there (include/tdm_burst_b200.h).
#pragma once
#include <dsp/processor.h>

#include "tdm_b200.h"
#include "tdm_burst_b200.h"

namespace dsp::b200 {

    class BurstSync : public Processor<uint8_t, uint8_t> {
        using base_type = Processor<uint8_t, uint8_t>;
    public:
        BurstSync() {}
        ~BurstSync();

        void init(stream<uint8_t>* in, int device = 0, bool detectTrainingSequences = false);

        int run() {
            int count = base_type::_in->read();
            if (count < 0) { return -1; }
            int outBytes = process(count, base_type::_in->readBuf, base_type::out.writeBuf);
            base_type::_in->flush();
            if (outBytes < 0) { return -1; }                     // CUDA / ABI failure: end the worker
            if (outBytes) {
                if (!base_type::out.swap(outBytes)) { return -1; }
            }
            return outBytes;
        }

        // returns the number of BYTES written to out (a multiple of sizeof(tdm_burst)), or -1 on failure
        int process(int count, const uint8_t* in, uint8_t* out);

        void reset();

        // enum rx_state of the receiver (phy/tetra_burst_sync.h:6-10) and the network mode's LED (src/main.cpp:471)
        int rxState() const { return state.state; }
        bool tsFound() const { return state.ts_found != 0; }
        unsigned long long burstsDelivered() const { return state.n_bursts; }
        const char* lastError() const;

    private:
        tdm_bsync* handle = nullptr;
        tdm_bsync_state state{};
        bool detect = false;
        static constexpr int kMaxBits = 1000000;                 // STREAM_BUFFER_SIZE
    };
}
