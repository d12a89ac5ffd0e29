// pi4dqpsk_b200.cpp -- see pi4dqpsk_b200.h.  Host-side glue only: every output byte comes out of libtdm_b200.so.
#include "pi4dqpsk_b200.h"

#include <stdio.h>
#include <string.h>
#include <algorithm>

namespace dsp::b200 {

    void SymbolFifo::push(const uint8_t* data, int nsym, float standarderr, bool sync) {
        {
            std::lock_guard<std::mutex> lck(mtx);
            if (stopped) { return; }
            q.insert(q.end(), data, data + (size_t)nsym * (size_t)width);
            // nobody reading (no block wired behind, or it is not started): keep the newest kMaxSymbols
            const size_t cap = kMaxSymbols * (size_t)width;
            if (q.size() > cap) { q.erase(q.begin(), q.begin() + (long)(q.size() - cap)); }
            lastErr = standarderr;
            lastSync = sync;
        }
        cv.notify_all();
    }
    bool SymbolFifo::pop(uint8_t* out, int nsym, float* standarderr, bool* sync) {
        const size_t need = (size_t)nsym * (size_t)width;
        std::unique_lock<std::mutex> lck(mtx);
        cv.wait(lck, [&] { return q.size() >= need || stopped; });
        if (q.size() < need) { return false; }
        std::copy(q.begin(), q.begin() + (long)need, out);
        q.erase(q.begin(), q.begin() + (long)need);
        if (standarderr) { *standarderr = lastErr; }
        if (sync) { *sync = lastSync; }
        return true;
    }
    void SymbolFifo::stop() {
        {
            std::lock_guard<std::mutex> lck(mtx);
            stopped = true;
        }
        cv.notify_all();
    }
    void SymbolFifo::restart() {
        std::lock_guard<std::mutex> lck(mtx);
        stopped = false;
        q.clear();
    }
    void SymbolFifo::resume() {
        std::lock_guard<std::mutex> lck(mtx);
        stopped = false;
    }

    namespace {
        thread_local PI4DQPSK* g_last_initialised = nullptr;
    }
    PI4DQPSK* PI4DQPSK::lastInitialised() { return g_last_initialised; }

    PI4DQPSK::~PI4DQPSK() {
        if (!base_type::_block_init) { return; }
        base_type::stop();
        dibitFifo.stop();
        bitFifo.stop();
        if (g_last_initialised == this) { g_last_initialised = nullptr; }
        tdm_destroy(handle);
        handle = nullptr;
    }

    // PI4DQPSK::init (src/dsp/pi4dqpsk.cpp:11-30): all tap/gain design happens inside tdm_create on the host,
    // from the same arguments.
    void PI4DQPSK::init(stream<complex_t>* in, double symbolrate, double samplerate, int rrcTapCount, double rrcBeta,
                        double agcRate, double costasBandwidth, double fllBandwidth, double omegaGain, double muGain,
                        double omegaRelLimit, int dev) {
        memset(&cfg, 0, sizeof(cfg));
        cfg.symbolrate = symbolrate;
        cfg.samplerate = samplerate;
        cfg.rrc_tap_count = rrcTapCount;
        cfg.rrc_beta = rrcBeta;
        cfg.agc_rate = agcRate;
        cfg.costas_bandwidth = costasBandwidth;
        cfg.fll_bandwidth = fllBandwidth;
        cfg.omega_gain = omegaGain;
        cfg.mu_gain = muGain;
        cfg.omega_rel_limit = omegaRelLimit;
        device = dev;
        if (handle) { tdm_destroy(handle); handle = nullptr; }
        // one channel per block instance, buffers as large as an SDR++ stream buffer
        const int rc = tdm_create(&cfg, 1, STREAM_BUFFER_SIZE, device, &handle);
        if (rc != TDM_OK || !handle) {
            // init() returns void in the reference: say why the block will not run instead of failing at the first buffer
            handle = nullptr;
            fprintf(stderr, "[tetra_demodulator b200] PI4DQPSK::init: tdm_create failed (%d): %s -- the block will not run\n", rc, tdm_last_error());
        } else {
            const int64_t s = tdm_max_symbols(handle, STREAM_BUFFER_SIZE);
            dibitBuf.resize((size_t)s);
            bitBuf.resize((size_t)(2 * s));
        }
        g_last_initialised = this;
        base_type::init(in);
    }

    const char* PI4DQPSK::lastError() const { return tdm_last_error(); }

    // The reference's setters (src/dsp/pi4dqpsk.cpp:31-118): the rate/RRC ones stop the worker, mutate, restart; the
    // coefficient ones only take ctrlMtx and write a float the worker reads.  Here every setter pauses the worker:
    // the handle is single-caller and the new design must not be swapped under a running tdm_process.  The EFFECT of
    // each setter is the reference's (tdm_set_params): e.g. setSymbolrate redesigns the RRC taps and restarts the
    // timing loop but leaves the band-edge filters alone.
    void PI4DQPSK::reconfigure(uint32_t what) {
        assert(base_type::_block_init);
        std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
        base_type::tempStop();
        if (handle) { tdm_set_params(handle, &cfg, what); }
        base_type::tempStart();
    }
    void PI4DQPSK::setSymbolrate(double symbolrate) { cfg.symbolrate = symbolrate; reconfigure(TDM_SET_RATES); }
    void PI4DQPSK::setSamplerate(double samplerate) { cfg.samplerate = samplerate; reconfigure(TDM_SET_RATES); }
    void PI4DQPSK::setRRCParams(int rrcTapCount, double rrcBeta) { cfg.rrc_tap_count = rrcTapCount; cfg.rrc_beta = rrcBeta; reconfigure(TDM_SET_RRC); }
    void PI4DQPSK::setRRCTapCount(int rrcTapCount) { setRRCParams(rrcTapCount, cfg.rrc_beta); }
    void PI4DQPSK::setRRCBeta(int rrcBeta) { setRRCParams(cfg.rrc_tap_count, rrcBeta); }
    void PI4DQPSK::setAGCRate(double agcRate) { cfg.agc_rate = agcRate; reconfigure(TDM_SET_AGC_RATE); }
    void PI4DQPSK::setCostasBandwidth(double bandwidth) { cfg.costas_bandwidth = bandwidth; reconfigure(TDM_SET_COSTAS_BW); }
    void PI4DQPSK::setFllBandwidth(double fllBandwidth) { cfg.fll_bandwidth = fllBandwidth; reconfigure(TDM_SET_FLL_BW); }
    void PI4DQPSK::setMMParams(double omegaGain, double muGain, double omegaRelLimit) {
        cfg.omega_gain = omegaGain; cfg.mu_gain = muGain; cfg.omega_rel_limit = omegaRelLimit; reconfigure(TDM_SET_TIMING_GAINS);
    }
    void PI4DQPSK::setOmegaGain(double omegaGain) { cfg.omega_gain = omegaGain; reconfigure(TDM_SET_TIMING_GAINS); }
    void PI4DQPSK::setMuGain(double muGain) { cfg.mu_gain = muGain; reconfigure(TDM_SET_TIMING_GAINS); }
    void PI4DQPSK::setOmegaRelLimit(double omegaRelLimit) { cfg.omega_rel_limit = omegaRelLimit; reconfigure(TDM_SET_TIMING_GAINS); }

    void PI4DQPSK::reset() {
        assert(base_type::_block_init);
        std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
        base_type::tempStop();
        if (handle) { tdm_reset(handle); }
        base_type::tempStart();
    }

    // PI4DQPSK::process (src/dsp/pi4dqpsk.cpp:132-140) + the two blocks that follow it in src/main.cpp:90-91,
    // one library call.  `in` and `out` are host buffers (SDR++ stream buffers): the library stages them.
    int PI4DQPSK::process(int count, const complex_t* in, complex_t* out) {
        if (!handle) { return -1; }
        const int64_t stride = tdm_max_symbols(handle, count);
        if ((size_t)stride > dibitBuf.size()) { dibitBuf.resize((size_t)stride); bitBuf.resize((size_t)(2 * stride)); }
        int rc = tdm_process(handle, reinterpret_cast<const float*>(in), count, count, reinterpret_cast<float*>(out),
                             dibitBuf.data(), bitBuf.data(), stride, &symCount,
                             TDM_OUT_SYMBOLS | TDM_OUT_DIBITS | TDM_OUT_BITS, TDM_MEM_HOST);
        if (rc != TDM_OK) { return -1; }
        if (symCount > 0) {
            tdm_metrics m;
            float se = 0;
            bool sy = false;
            if (tdm_get_metrics(handle, &m, 1) == TDM_OK) { se = m.standarderr; sy = m.sync != 0; }
            dibitFifo.push(dibitBuf.data(), symCount, se, sy);
            bitFifo.push(bitBuf.data(), symCount, se, sy);
        }
        return symCount;
    }

    // DQPSKSymbolExtractor::process (src/dsp/dqpsk_sym_extr.cpp:4-55): same count in, same count out.  The symbols in
    // `in` are the fused launch's own output; their decisions were taken in the same launch and wait in the FIFO.
    int DQPSKSymbolExtractor::process(int count, const complex_t* in, uint8_t* out) {
        (void)in;
        if (!src || !src->dibitFifo.pop(out, count, &standarderr, &sync)) { return -1; }
        return count;
    }

    // BitUnpacker::process (src/dsp/bit_unpacker.cpp:4-10): returns count*2.
    int BitUnpacker::process(int count, const uint8_t* in, uint8_t* out) {
        (void)in;
        if (!src || !src->bitFifo.pop(out, count, nullptr, nullptr)) { return -1; }
        return count * 2;
    }
}
