// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/multirate/polyphase_bank.h>; restates SURVEY.md Appendix A.6.
#pragma once
#include "../processor.h"

namespace dsp::multirate {
    template <class T>
    struct PolyphaseBank {
        int phaseCount = 0;
        int tapsPerPhase = 0;
        T** phases = nullptr;
    };

    // phases[(P-1) - (i % P)][i / P] = taps[i]   (phase order reversed)
    template <class T>
    inline PolyphaseBank<T> buildPolyphaseBank(int phaseCount, tap<T>& taps) {
        PolyphaseBank<T> pb;
        pb.phaseCount = phaseCount;
        pb.phases = buffer::alloc<T*>(phaseCount);
        pb.tapsPerPhase = (taps.size + phaseCount - 1) / phaseCount;
        for (int i = 0; i < phaseCount; i++) {
            pb.phases[i] = buffer::alloc<T>(pb.tapsPerPhase);
        }
        int totTapCount = phaseCount * pb.tapsPerPhase;
        for (int i = 0; i < totTapCount; i++) {
            pb.phases[(phaseCount - 1) - (i % phaseCount)][i / phaseCount] = (i < taps.size) ? taps.taps[i] : 0;
        }
        return pb;
    }

    template <class T>
    inline void freePolyphaseBank(PolyphaseBank<T>& bank) {
        if (!bank.phases) { return; }
        for (int i = 0; i < bank.phaseCount; i++) {
            if (bank.phases[i]) { buffer::free(bank.phases[i]); }
        }
        buffer::free(bank.phases);
        bank.phases = nullptr;
        bank.phaseCount = 0;
        bank.tapsPerPhase = 0;
    }
}
