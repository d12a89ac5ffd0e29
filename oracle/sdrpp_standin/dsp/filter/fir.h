// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/filter/fir.h>; restates SURVEY.md Appendix A.4.
#pragma once
#include <type_traits>
#include "../processor.h"
#include <volk/volk.h>

namespace dsp::filter {
    template <class D, class T>
    class FIR : public Processor<D, D> {
        using base_type = Processor<D, D>;
    public:
        FIR() {}
        ~FIR() {
            if (!base_type::_block_init) { return; }
            base_type::stop();
            buffer::free(buffer);
        }

        virtual void init(stream<D>* in, tap<T>& taps) {
            _taps = taps;
            buffer = buffer::alloc<D>(STREAM_BUFFER_SIZE + 64000);
            bufStart = &buffer[_taps.size - 1];
            buffer::clear<D>(buffer, _taps.size - 1);
            base_type::init(in);
        }

        virtual void setTaps(tap<T>& taps) {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            base_type::tempStop();
            int oldTC = _taps.size;
            _taps = taps;
            // Keep as much history as the new length allows, newest samples last.
            if (_taps.size < oldTC) {
                memmove(buffer, &buffer[oldTC - _taps.size], (_taps.size - 1) * sizeof(D));
            }
            else if (_taps.size > oldTC) {
                memmove(&buffer[_taps.size - oldTC], buffer, (oldTC - 1) * sizeof(D));
                buffer::clear<D>(buffer, _taps.size - oldTC);
            }
            bufStart = &buffer[_taps.size - 1];
            base_type::tempStart();
        }

        virtual void reset() {
            assert(base_type::_block_init);
            std::lock_guard<std::recursive_mutex> lck(base_type::ctrlMtx);
            base_type::tempStop();
            buffer::clear<D>(buffer, _taps.size - 1);
            base_type::tempStart();
        }

        // out[i] = sum_k buffer[i+k] * taps[k]  (no reversal, no conjugate)
        inline int process(int count, const D* in, D* out) {
            memcpy(bufStart, in, count * sizeof(D));
            for (int i = 0; i < count; i++) {
                if constexpr (std::is_same_v<D, complex_t> && std::is_same_v<T, float>) {
                    volk_32fc_32f_dot_prod_32fc((lv_32fc_t*)&out[i], (lv_32fc_t*)&buffer[i], _taps.taps, _taps.size);
                }
                else if constexpr (std::is_same_v<D, complex_t> && std::is_same_v<T, complex_t>) {
                    volk_32fc_x2_dot_prod_32fc((lv_32fc_t*)&out[i], (lv_32fc_t*)&buffer[i], (lv_32fc_t*)_taps.taps, _taps.size);
                }
                else {
                    volk_32f_x2_dot_prod_32f((float*)&out[i], (float*)&buffer[i], (float*)_taps.taps, _taps.size);
                }
            }
            memmove(buffer, &buffer[count], (_taps.size - 1) * sizeof(D));
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
        tap<T> _taps;
        D* buffer = nullptr;
        D* bufStart = nullptr;
    };
}
