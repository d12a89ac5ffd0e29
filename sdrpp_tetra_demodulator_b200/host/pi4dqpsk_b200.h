// pi4dqpsk_b200.h -- SDR++ blocks that stand in for the reference's demodulation chain and run it on a B200.
//
// Drop-in surface.  The reference wires (src/main.cpp:84-91)
//
//     mainDemodulator (dsp::demod::PI4DQPSK)  ->  Splitter -> { constellation sink, demodStream }
//     demodStream -> symbolExtractor (dsp::DQPSKSymbolExtractor) -> bitsUnpacker (dsp::BitUnpacker) -> decoder / UDP
//
// This header provides the same three class shapes in namespace dsp::b200 -- same base classes
// (dsp::Processor<I,O>), same init()/process()/run()/reset()/setter signatures, same public members
// (`out`, `sync`, `standarderr`) -- so src/main.cpp changes three type names and nothing else:
//
//     dsp::b200::PI4DQPSK              : Processor<complex_t, complex_t>   (src/dsp/pi4dqpsk.h:27-81)
//     dsp::b200::DQPSKSymbolExtractor  : Processor<complex_t, uint8_t>     (src/dsp/dqpsk_sym_extr.h:19-46)
//     dsp::b200::BitUnpacker           : Processor<uint8_t, uint8_t>       (src/dsp/bit_unpacker.h:16-34)
//
// The GPU kernel is fused (one launch = demodulate + slice + decode + unpack), so PI4DQPSK keeps, next to the
// complex symbols it puts on `out`, the dibits and bits of the same call in a side queue keyed by that call's
// symbol count; the two downstream blocks pop from it instead of recomputing.  Their run() loops, stream
// hand-offs and stop behaviour are the reference's own (same code shape as src/dsp/pi4dqpsk.h:38-50).
//
// Everything below the class surface is the C ABI of libtdm_b200.so (include/tdm_b200.h).  If the library
// reports an error, run() returns -1 and the worker thread ends -- the reference's only failure signal
// (SURVEY.md 8b "Errors").  Built against SDR++ core headers in the plugin tree; against the stand-in headers
// in oracle/sdrpp_standin for the tests here.
#pragma once
#include <dsp/processor.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

#include "tdm_b200.h"

namespace dsp::b200 {

    // One fused call's slicer outputs, handed from PI4DQPSK to the extractor/unpacker blocks downstream.
    struct FusedBatch {
        int nsym = 0;
        std::vector<uint8_t> dibits;   // nsym
        std::vector<uint8_t> bits;     // 2*nsym
        float standarderr = 0;
        bool sync = false;
    };

    class FusedQueue {
    public:
        void push(FusedBatch&& b);
        // blocks until a batch is available or stop() was called; false on stop
        bool pop(FusedBatch& out);
        void stop();
        void restart();
    private:
        std::mutex mtx;
        std::condition_variable cv;
        std::deque<FusedBatch> q;
        bool stopped = false;
    };

    class PI4DQPSK : public Processor<complex_t, complex_t> {
        using base_type = Processor<complex_t, complex_t>;
    public:
        PI4DQPSK() {}
        PI4DQPSK(stream<complex_t>* in, double symbolrate, double samplerate, int rrcTapCount, double rrcBeta,
                 double agcRate, double costasBandwidth, double fllBandwidth, double omegaGain, double muGain,
                 double omegaRelLimit = 0.01) {
            // the reference's convenience constructor drops omegaRelLimit (src/dsp/pi4dqpsk.h:32); kept as is
            init(in, symbolrate, samplerate, rrcTapCount, rrcBeta, agcRate, costasBandwidth, fllBandwidth, omegaGain, muGain);
        }
        ~PI4DQPSK();

        // same argument list as the reference (src/dsp/pi4dqpsk.h:36); `device` selects the CUDA device
        void init(stream<complex_t>* in, double symbolrate, double samplerate, int rrcTapCount, double rrcBeta,
                  double agcRate, double costasBandwidth, double fllBandwidth, double omegaGain, double muGain,
                  double omegaRelLimit = 0.01, int device = 0);

        int run() {
            int count = base_type::_in->read();
            if (count < 0) { return -1; }
            int outCount = process(count, base_type::_in->readBuf, base_type::out.writeBuf);
            base_type::_in->flush();
            if (outCount < 0) { return -1; }                   // CUDA / ABI failure: end the worker
            if (outCount) {
                if (!base_type::out.swap(outCount)) { return -1; }
            }
            return outCount;
        }

        void setSymbolrate(double symbolrate);
        void setSamplerate(double samplerate);
        void setRRCParams(int rrcTapCount, double rrcBeta);
        void setRRCTapCount(int rrcTapCount);
        void setRRCBeta(int rrcBeta);                           // int, like the reference (truncates, [A.9])
        void setAGCRate(double agcRate);
        void setCostasBandwidth(double bandwidth);
        void setFllBandwidth(double fllBandwidth);
        void setMMParams(double omegaGain, double muGain, double omegaRelLimit = 0.01);
        void setOmegaGain(double omegaGain);
        void setMuGain(double muGain);
        void setOmegaRelLimit(double omegaRelLimit);

        void reset();

        // returns the number of symbols written to `out`, or -1 if the GPU call failed
        int process(int count, const complex_t* in, complex_t* out);

        // side channel for the fused slicer results (consumed by DQPSKSymbolExtractor below)
        FusedQueue fused;
        const char* lastError() const;

    protected:
        void reconfigure();
        tdm_config cfg{};
        tdm_handle* handle = nullptr;
        int device = 0;
        std::vector<uint8_t> dibitBuf, bitBuf;
        int32_t symCount = 0;
    };

    // Symbol mapper + differential decoder: hands out what the fused kernel already computed for the symbols
    // it is given.  `sync` / `standarderr` are the public members the GUI reads (src/main.cpp:211-217).
    class DQPSKSymbolExtractor : public Processor<complex_t, uint8_t> {
        using base_type = Processor<complex_t, uint8_t>;
    public:
        void init(stream<complex_t>* in, PI4DQPSK* source) {
            src = source;
            base_type::init(in);
        }
        int run() {
            int count = base_type::_in->read();
            if (count < 0) { return -1; }
            int outCount = process(count, base_type::_in->readBuf, base_type::out.writeBuf);
            base_type::_in->flush();
            if (outCount < 0) { return -1; }
            if (outCount) {
                if (!base_type::out.swap(outCount)) { return -1; }
            }
            return outCount;
        }
        int process(int count, const complex_t* in, uint8_t* out);

        bool sync = false;
        float standarderr = 0;
        // bits of the batch most recently handed out, for the BitUnpacker that follows
        FusedQueue unpacked;

    private:
        PI4DQPSK* src = nullptr;
    };

    // dibit/byte -> 2 x bit/byte, MSB first (src/dsp/bit_unpacker.cpp:6-7): the fused kernel wrote them already.
    class BitUnpacker : public Processor<uint8_t, uint8_t> {
        using base_type = Processor<uint8_t, uint8_t>;
    public:
        void init(stream<uint8_t>* in, DQPSKSymbolExtractor* source) {
            src = source;
            base_type::init(in);
        }
        int run() {
            int count = base_type::_in->read();
            if (count < 0) { return -1; }
            int outCount = process(count, base_type::_in->readBuf, base_type::out.writeBuf);
            base_type::_in->flush();
            if (outCount < 0) { return -1; }
            if (outCount) {
                if (!base_type::out.swap(outCount)) { return -1; }
            }
            return outCount;
        }
        int process(int count, const uint8_t* in, uint8_t* out);

    private:
        DQPSKSymbolExtractor* src = nullptr;
    };
}
