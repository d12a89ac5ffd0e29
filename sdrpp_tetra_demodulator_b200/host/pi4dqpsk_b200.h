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
// complex symbols it puts on `out`, the dibits and bits of the same symbols in two FIFOs indexed by CUMULATIVE
// SYMBOL NUMBER; the two downstream blocks take as many entries as their input buffer holds symbols, whatever the
// buffer boundaries are (a Reshaper or any other re-chunking block in between changes nothing).  Their run()
// loops, stream hand-offs and stop behaviour are the reference's own (same code shape as src/dsp/pi4dqpsk.h:38-50).
// PI4DQPSK::start() empties both FIFOs (buffers dropped by a stop()/start() cycle must not leave stale entries behind:
// src/main.cpp enable()/disable() stop and start all blocks together; like in the reference, data in flight around a
// stop is lost); a downstream block's stop() interrupts its own wait on the FIFO; a FIFO nobody reads is bounded (the
// oldest entries are dropped).  Setters (tempStop/tempStart) leave the FIFOs alone.
//
// The downstream blocks keep the reference's init() signatures: init(in) binds to the PI4DQPSK most recently
// initialised on the calling thread (src/main.cpp:84-91 initialises the three blocks of an instance back to back);
// init(in, source) binds explicitly.
//
// Everything below the class surface is the C ABI of libtdm_b200.so (include/tdm_b200.h).  The reference's blocks
// have no error codes (SURVEY.md 8b "Errors"): if the library cannot be brought up, init() says so on stderr,
// ok() is false and run() returns -1, which ends the worker thread -- the reference's only failure signal.
// Built against SDR++ core headers in the plugin tree; against the stand-in headers in oracle/sdrpp_standin for
// the tests here.
#pragma once
#include <dsp/processor.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <vector>

#include "tdm_b200.h"

namespace dsp::b200 {

    // Slicer outputs of the fused launches as a stream: `width` bytes per symbol (1: dibits, 2: bits), first symbol
    // number `head`.  One writer (PI4DQPSK::process), one reader (the block downstream).
    class SymbolFifo {
    public:
        explicit SymbolFifo(int width_) : width(width_) {}
        void push(const uint8_t* data, int nsym, float standarderr, bool sync);
        // the next nsym symbols' bytes; blocks until they are there or stop() was called (false)
        bool pop(uint8_t* out, int nsym, float* standarderr, bool* sync);
        void stop();
        void restart();                 // empty, accept data again (PI4DQPSK::start)
        void resume();                  // accept data again, keep what is there (the reading block's start)
        static constexpr size_t kMaxSymbols = 8u * 1000000u;   // 8 SDR++ stream buffers
    private:
        const int width;
        std::mutex mtx;
        std::condition_variable cv;
        std::deque<uint8_t> q;
        float lastErr = 0;
        bool lastSync = false;
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

        // side channels for the fused slicer results (consumed by DQPSKSymbolExtractor / BitUnpacker below)
        SymbolFifo dibitFifo{ 1 };
        SymbolFifo bitFifo{ 2 };
        const char* lastError() const;
        bool ok() const { return handle != nullptr; }          // false: tdm_create failed in init(), run() returns -1
        static PI4DQPSK* lastInitialised();                     // on the calling thread

        void start() override {
            dibitFifo.restart();
            bitFifo.restart();
            base_type::start();
        }

    protected:
        void reconfigure(uint32_t what);
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
        void init(stream<complex_t>* in) { init(in, PI4DQPSK::lastInitialised()); }      // src/dsp/dqpsk_sym_extr.h:27
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
        PI4DQPSK* source() const { return src; }

        bool sync = false;
        float standarderr = 0;

    protected:
        void doStart() override {
            if (src) { src->dibitFifo.resume(); }
            base_type::doStart();
        }
        void doStop() override {
            if (src) { src->dibitFifo.stop(); }            // a worker waiting for dibits must see the stop
            base_type::doStop();
        }

    private:
        PI4DQPSK* src = nullptr;
    };

    // dibit/byte -> 2 x bit/byte, MSB first (src/dsp/bit_unpacker.cpp:6-7): the fused kernel wrote them already.
    class BitUnpacker : public Processor<uint8_t, uint8_t> {
        using base_type = Processor<uint8_t, uint8_t>;
    public:
        void init(stream<uint8_t>* in) { init(in, PI4DQPSK::lastInitialised()); }         // src/dsp/bit_unpacker.h:24
        void init(stream<uint8_t>* in, DQPSKSymbolExtractor* source) { init(in, source ? source->source() : nullptr); }
        void init(stream<uint8_t>* in, PI4DQPSK* source) {
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

    protected:
        void doStart() override {
            if (src) { src->bitFifo.resume(); }
            base_type::doStart();
        }
        void doStop() override {
            if (src) { src->bitFifo.stop(); }
            base_type::doStop();
        }

    private:
        PI4DQPSK* src = nullptr;
    };
}
