// TEST INFRASTRUCTURE (oracle side) -- stand-in for SDR++ core
// <dsp/processor.h>, <dsp/block.h>, <dsp/stream.h>, <dsp/buffer/buffer.h>,
// <dsp/taps/tap.h>.  From-scratch restatement of SURVEY.md Appendix A.7; see
// dsp/types.h in this directory for why it exists.  Not product code.
#pragma once
#include <assert.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "types.h"

#define STREAM_BUFFER_SIZE 1000000

namespace dsp {
    namespace buffer {
        template <class T>
        inline T* alloc(int count) {
            // Upstream hands out uninitialised (volk_malloc) memory; the oracles
            // DEFINE never-written history as zeros, so zero it.
            // calloc: large blocks come back as lazily-zeroed pages, so the many
            // 8 MB stream buffers the block classes own cost no resident memory.
            void* p = calloc((size_t)std::max(count, 1), sizeof(T));
            if (!p) { abort(); }
            return (T*)p;
        }
        template <class T>
        inline void clear(T* buffer, int count, int offset = 0) {
            memset(&buffer[offset], 0, sizeof(T) * (size_t)count);
        }
        inline void free(void* buffer) { ::free(buffer); }
    }

    template <class T>
    struct tap {
        T* taps = nullptr;
        int size = 0;
    };

    namespace taps {
        template <class T>
        inline tap<T> alloc(int count) {
            tap<T> t;
            t.size = count;
            t.taps = buffer::alloc<T>(count);
            return t;
        }
        template <class T>
        inline void free(tap<T>& t) {
            if (t.taps) { buffer::free(t.taps); }
            t.taps = nullptr;
            t.size = 0;
        }
    }

    class untyped_stream {
    public:
        virtual ~untyped_stream() {}
        virtual bool swap(int size) { return false; }
        virtual int read() { return -1; }
        virtual void flush() {}
        virtual void stopWriter() {}
        virtual void clearWriteStop() {}
        virtual void stopReader() {}
        virtual void clearReadStop() {}
    };

    // Double-buffered hand-off between two block threads (A.7).
    template <class T>
    class stream : public untyped_stream {
    public:
        stream() {
            writeBuf = buffer::alloc<T>(STREAM_BUFFER_SIZE);
            readBuf = buffer::alloc<T>(STREAM_BUFFER_SIZE);
        }
        virtual ~stream() { free(); }

        virtual bool swap(int size) {
            {
                std::unique_lock<std::mutex> lck(swapMtx);
                swapCV.wait(lck, [this] { return canSwap || writerStop; });
                if (writerStop) { return false; }
                dataSize = size;
                std::swap(writeBuf, readBuf);
                canSwap = false;
            }
            {
                std::lock_guard<std::mutex> lck(rdyMtx);
                dataReady = true;
            }
            rdyCV.notify_all();
            return true;
        }

        virtual int read() {
            std::unique_lock<std::mutex> lck(rdyMtx);
            rdyCV.wait(lck, [this] { return dataReady || readerStop; });
            return readerStop ? -1 : dataSize;
        }

        virtual void flush() {
            {
                std::lock_guard<std::mutex> lck(rdyMtx);
                dataReady = false;
            }
            {
                std::lock_guard<std::mutex> lck(swapMtx);
                canSwap = true;
            }
            swapCV.notify_all();
        }

        virtual void stopWriter() {
            {
                std::lock_guard<std::mutex> lck(swapMtx);
                writerStop = true;
            }
            swapCV.notify_all();
        }
        virtual void clearWriteStop() { writerStop = false; }
        virtual void stopReader() {
            {
                std::lock_guard<std::mutex> lck(rdyMtx);
                readerStop = true;
            }
            rdyCV.notify_all();
        }
        virtual void clearReadStop() { readerStop = false; }

        void free() {
            if (writeBuf) { buffer::free(writeBuf); }
            if (readBuf) { buffer::free(readBuf); }
            writeBuf = nullptr;
            readBuf = nullptr;
        }

        T* writeBuf = nullptr;
        T* readBuf = nullptr;

    private:
        std::mutex swapMtx;
        std::condition_variable swapCV;
        bool canSwap = true;

        std::mutex rdyMtx;
        std::condition_variable rdyCV;
        bool dataReady = false;

        bool readerStop = false;
        bool writerStop = false;
        int dataSize = 0;
    };

    class block {
    public:
        virtual ~block() {}

        virtual void start() {
            assert(_block_init);
            std::lock_guard<std::recursive_mutex> lck(ctrlMtx);
            if (running) { return; }
            running = true;
            doStart();
        }
        virtual void stop() {
            assert(_block_init);
            std::lock_guard<std::recursive_mutex> lck(ctrlMtx);
            if (!running) { return; }
            doStop();
            running = false;
        }
        void tempStart() {
            assert(_block_init);
            if (!tempStopDepth || --tempStopDepth) { return; }
            if (tempStopped) {
                doStart();
                tempStopped = false;
            }
        }
        void tempStop() {
            assert(_block_init);
            if (tempStopDepth++) { return; }
            if (running && !tempStopped) {
                doStop();
                tempStopped = true;
            }
        }
        virtual int run() = 0;

    protected:
        void workerLoop() {
            while (run() >= 0) {}
        }
        void registerInput(untyped_stream* s) { inputs.push_back(s); }
        void unregisterInput(untyped_stream* s) {
            inputs.erase(std::remove(inputs.begin(), inputs.end(), s), inputs.end());
        }
        void registerOutput(untyped_stream* s) { outputs.push_back(s); }
        void unregisterOutput(untyped_stream* s) {
            outputs.erase(std::remove(outputs.begin(), outputs.end(), s), outputs.end());
        }
        virtual void doStart() { workerThread = std::thread(&block::workerLoop, this); }
        virtual void doStop() {
            for (auto& in : inputs) { if (in) { in->stopReader(); } }
            for (auto& out : outputs) { out->stopWriter(); }
            if (workerThread.joinable()) { workerThread.join(); }
            for (auto& in : inputs) { if (in) { in->clearReadStop(); } }
            for (auto& out : outputs) { out->clearWriteStop(); }
        }

        bool _block_init = false;
        std::recursive_mutex ctrlMtx;
        std::vector<untyped_stream*> inputs;
        std::vector<untyped_stream*> outputs;
        bool running = false;
        bool tempStopped = false;
        int tempStopDepth = 0;
        std::thread workerThread;
    };

    template <class I, class O>
    class Processor : public block {
    public:
        Processor() {}
        Processor(stream<I>* in) { init(in); }
        virtual ~Processor() {}

        virtual void init(stream<I>* in) {
            _in = in;
            registerInput(_in);
            registerOutput(&out);
            _block_init = true;
        }
        virtual void setInput(stream<I>* in) {
            assert(_block_init);
            std::lock_guard<std::recursive_mutex> lck(ctrlMtx);
            tempStop();
            unregisterInput(_in);
            _in = in;
            registerInput(_in);
            tempStart();
        }
        virtual int run() = 0;

        stream<O> out;

    protected:
        stream<I>* _in = nullptr;
    };
}
