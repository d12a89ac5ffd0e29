// burst_sync_b200.cpp -- see burst_sync_b200.h.  Thin: argument plumbing around tdm_bsync_in.
#include "burst_sync_b200.h"

#include <algorithm>

namespace dsp::b200 {

    BurstSync::~BurstSync() {
        if (!base_type::_block_init) { return; }
        base_type::stop();
        if (handle) { tdm_bsync_destroy(handle); handle = nullptr; }
    }

    void BurstSync::init(stream<uint8_t>* in, int device, bool detectTrainingSequences) {
        detect = detectTrainingSequences;
        if (tdm_bsync_create(1, kMaxBits, device, &handle) != TDM_OK) { handle = nullptr; }
        base_type::init(in);
    }

    void BurstSync::reset() {
        std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
        base_type::tempStop();
        if (handle) { tdm_bsync_reset(handle); }
        state = tdm_bsync_state{};
        base_type::tempStart();
    }

    int BurstSync::process(int count, const uint8_t* in, uint8_t* out) {
        if (!handle || count < 0 || count > kMaxBits) { return -1; }
        if (count == 0) { return 0; }
        const int callBits = std::min(count, TDM_BSYNC_MAX_CALL_BITS);
        const int maxBursts = (count + callBits - 1) / callBits;   // at most one burst per emulated call
        int32_t n = 0;
        if (tdm_bsync_in(handle, in, count, nullptr, count, TDM_BSYNC_IN_BITS, callBits, reinterpret_cast<tdm_burst*>(out), maxBursts, &n,
                         detect ? 1 : 0, TDM_MEM_HOST) != TDM_OK) {
            return -1;
        }
        if (tdm_bsync_get_state(handle, &state, 1) != TDM_OK) { return -1; }
        return n * (int)sizeof(tdm_burst);
    }

    const char* BurstSync::lastError() const { return tdm_last_error(); }
}
