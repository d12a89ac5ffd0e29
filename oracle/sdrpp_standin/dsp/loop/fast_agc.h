// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/loop/fast_agc.h>; restates SURVEY.md Appendix A.3.
#pragma once
#include <type_traits>
#include "../processor.h"

namespace dsp::loop {
    template <class T>
    class FastAGC : public Processor<T, T> {
        using base_type = Processor<T, T>;
    public:
        FastAGC() {}
        void init(stream<T>* in, double setPoint, double maxGain, double rate, double initGain = 1.0) {
            _setPoint = setPoint;
            _maxGain = maxGain;
            _rate = rate;
            _initGain = initGain;
            _gain = _initGain;
            base_type::init(in);
        }
        void setSetPoint(double setPoint) { _setPoint = setPoint; }
        void setMaxGain(double maxGain) { _maxGain = maxGain; }
        void setRate(double rate) { _rate = rate; }
        void setInitGain(double initGain) { _initGain = initGain; }
        void setGain(double gain) { _gain = gain; }
        float getGain() const { return _gain; }
        void reset() { _gain = _initGain; }

        inline int process(int count, T* in, T* out) {
            for (int i = 0; i < count; i++) {
                out[i] = in[i] * _gain;
                float amp;
                if constexpr (std::is_same_v<T, float>) { amp = fabsf(out[i]); }
                else { amp = out[i].amplitude(); }
                _gain += (_setPoint - amp) * _rate;
                if (_gain > _maxGain) { _gain = _maxGain; }
            }
            return count;
        }

        int run() {
            int count = base_type::_in->read();
            if (count < 0) { return -1; }
            process(count, base_type::_in->readBuf, base_type::out.writeBuf);
            base_type::_in->flush();
            if (!base_type::out.swap(count)) { return -1; }
            return count;
        }

    protected:
        float _gain;
        float _setPoint;
        float _rate;
        float _maxGain;
        float _initGain;
    };
}
