// pi4dqpsk_b200.cpp -- see pi4dqpsk_b200.h.  Host-side glue only: every output byte comes out of libtdm_b200.so.
#include "pi4dqpsk_b200.h"

#include <string.h>

namespace dsp::b200 {

    void FusedQueue::push(FusedBatch&& b) {
        {
            std::lock_guard<std::mutex> lck(mtx);
            q.emplace_back(std::move(b));
        }
        cv.notify_all();
    }
    bool FusedQueue::pop(FusedBatch& out) {
        std::unique_lock<std::mutex> lck(mtx);
        cv.wait(lck, [this] { return !q.empty() || stopped; });
        if (q.empty()) { return false; }
        out = std::move(q.front());
        q.pop_front();
        return true;
    }
    void FusedQueue::stop() {
        {
            std::lock_guard<std::mutex> lck(mtx);
            stopped = true;
        }
        cv.notify_all();
    }
    void FusedQueue::restart() {
        std::lock_guard<std::mutex> lck(mtx);
        stopped = false;
        q.clear();
    }

    PI4DQPSK::~PI4DQPSK() {
        if (!base_type::_block_init) { return; }
        base_type::stop();
        fused.stop();
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
        tdm_create(&cfg, 1, STREAM_BUFFER_SIZE, device, &handle);
        if (handle) {
            const int64_t s = tdm_max_symbols(handle, STREAM_BUFFER_SIZE);
            dibitBuf.resize((size_t)s);
            bitBuf.resize((size_t)(2 * s));
        }
        base_type::init(in);
    }

    const char* PI4DQPSK::lastError() const { return tdm_last_error(); }

    // the reference's setters stop the worker, mutate, restart (src/dsp/pi4dqpsk.cpp:32-118)
    void PI4DQPSK::reconfigure() {
        assert(base_type::_block_init);
        std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
        base_type::tempStop();
        if (handle) { tdm_set_config(handle, &cfg); }
        base_type::tempStart();
    }
    void PI4DQPSK::setSymbolrate(double symbolrate) { cfg.symbolrate = symbolrate; reconfigure(); }
    void PI4DQPSK::setSamplerate(double samplerate) { cfg.samplerate = samplerate; reconfigure(); }
    void PI4DQPSK::setRRCParams(int rrcTapCount, double rrcBeta) { cfg.rrc_tap_count = rrcTapCount; cfg.rrc_beta = rrcBeta; reconfigure(); }
    void PI4DQPSK::setRRCTapCount(int rrcTapCount) { setRRCParams(rrcTapCount, cfg.rrc_beta); }
    void PI4DQPSK::setRRCBeta(int rrcBeta) { setRRCParams(cfg.rrc_tap_count, rrcBeta); }
    void PI4DQPSK::setAGCRate(double agcRate) { cfg.agc_rate = agcRate; reconfigure(); }
    void PI4DQPSK::setCostasBandwidth(double bandwidth) { cfg.costas_bandwidth = bandwidth; reconfigure(); }
    void PI4DQPSK::setFllBandwidth(double fllBandwidth) { cfg.fll_bandwidth = fllBandwidth; reconfigure(); }
    void PI4DQPSK::setMMParams(double omegaGain, double muGain, double omegaRelLimit) {
        cfg.omega_gain = omegaGain; cfg.mu_gain = muGain; cfg.omega_rel_limit = omegaRelLimit; reconfigure();
    }
    void PI4DQPSK::setOmegaGain(double omegaGain) { cfg.omega_gain = omegaGain; reconfigure(); }
    void PI4DQPSK::setMuGain(double muGain) { cfg.mu_gain = muGain; reconfigure(); }
    void PI4DQPSK::setOmegaRelLimit(double omegaRelLimit) { cfg.omega_rel_limit = omegaRelLimit; reconfigure(); }

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
            FusedBatch b;
            b.nsym = symCount;
            b.dibits.assign(dibitBuf.begin(), dibitBuf.begin() + symCount);
            b.bits.assign(bitBuf.begin(), bitBuf.begin() + 2 * (size_t)symCount);
            tdm_metrics m;
            if (tdm_get_metrics(handle, &m, 1) == TDM_OK) { b.standarderr = m.standarderr; b.sync = m.sync != 0; }
            fused.push(std::move(b));
        }
        return symCount;
    }

    // DQPSKSymbolExtractor::process (src/dsp/dqpsk_sym_extr.cpp:4-55): same count in, same count out.
    int DQPSKSymbolExtractor::process(int count, const complex_t* in, uint8_t* out) {
        (void)in;
        FusedBatch b;
        if (!src || !src->fused.pop(b) || b.nsym != count) { return -1; }
        memcpy(out, b.dibits.data(), (size_t)count);
        sync = b.sync;
        standarderr = b.standarderr;
        unpacked.push(std::move(b));
        return count;
    }

    // BitUnpacker::process (src/dsp/bit_unpacker.cpp:4-10): returns count*2.
    int BitUnpacker::process(int count, const uint8_t* in, uint8_t* out) {
        (void)in;
        FusedBatch b;
        if (!src || !src->unpacked.pop(b) || b.nsym != count) { return -1; }
        memcpy(out, b.bits.data(), 2 * (size_t)count);
        return count * 2;
    }
}
