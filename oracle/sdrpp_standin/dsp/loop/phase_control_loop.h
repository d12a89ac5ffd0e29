// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/loop/phase_control_loop.h>; restates SURVEY.md Appendix A.2.
#pragma once
#include <math.h>
#include "../types.h"

namespace dsp::loop {
    template <class T, bool CLAMP_PHASE = true>
    class PhaseControlLoop {
    public:
        PhaseControlLoop() {}
        PhaseControlLoop(T alpha, T beta, T phase, T minPhase, T maxPhase, T freq, T minFreq, T maxFreq) {
            init(alpha, beta, phase, minPhase, maxPhase, freq, minFreq, maxFreq);
        }

        void init(T alpha, T beta, T phase, T minPhase, T maxPhase, T freq, T minFreq, T maxFreq) {
            _alpha = alpha;
            _beta = beta;
            this->phase = phase;
            _minPhase = minPhase;
            _maxPhase = maxPhase;
            this->freq = freq;
            _minFreq = minFreq;
            _maxFreq = maxFreq;
            phaseDelta = _maxPhase - _minPhase;
        }

        // A.2: d = sqrt(2)/2; den = 1 + 2 d bw + bw^2; alpha = 4 d bw/den; beta = 4 bw^2/den
        static inline void criticallyDamped(T bandwidth, T& alpha, T& beta) {
            T dampningFactor = sqrt(2.0) / 2.0;
            T denominator = (1.0 + 2.0 * dampningFactor * bandwidth + bandwidth * bandwidth);
            alpha = (4 * dampningFactor * bandwidth) / denominator;
            beta = (4 * bandwidth * bandwidth) / denominator;
        }

        void setPhaseLimits(T minPhase, T maxPhase) {
            _minPhase = minPhase;
            _maxPhase = maxPhase;
            phaseDelta = _maxPhase - _minPhase;
            clampPhase();
        }
        void setFreqLimits(T minFreq, T maxFreq) {
            _minFreq = minFreq;
            _maxFreq = maxFreq;
            clampFreq();
        }
        void setCoefficients(T alpha, T beta) {
            _alpha = alpha;
            _beta = beta;
        }

        inline void advance(T error) {
            freq += _beta * error;
            clampFreq();
            phase += freq + (_alpha * error);
            if constexpr (CLAMP_PHASE) { clampPhase(); }
        }
        inline T advancePhase() {
            phase += freq;
            if constexpr (CLAMP_PHASE) { clampPhase(); }
            return phase;
        }

        T freq;
        T phase;

    protected:
        inline void clampFreq() {
            if (freq > _maxFreq) { freq = _maxFreq; }
            else if (freq < _minFreq) { freq = _minFreq; }
        }
        inline void clampPhase() {
            while (phase > _maxPhase) { phase -= phaseDelta; }
            while (phase < _minPhase) { phase += phaseDelta; }
        }

        T _alpha;
        T _beta;
        T _minPhase;
        T _maxPhase;
        T _minFreq;
        T _maxFreq;
        T phaseDelta;
    };
}
