// tdm_api.cu -- the extern "C" ABI of libtdm_b200.so (declared in include/tdm_b200.h).
//
// Thin by design: argument checking, device-memory ownership, host<->device staging for
// callers that hand in host buffers, and kernel launches.  There is NO CPU fallback: if
// no usable sm_100 device is present tdm_create fails with TDM_ERR_NO_DEVICE, and nothing
// in this library computes demodulator outputs on the host.
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "tdm_b200.h"
#include "tdm_kernels.cuh"
#include "tdm_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define TDM_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) { return fail(TDM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); } \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) { cudaSetDevice(dev); }
        else { prev = -1; }
    }
    ~DeviceGuard() {
        if (prev >= 0) { cudaSetDevice(prev); }
    }
};

}  // namespace

int tdm_internal_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

struct tdm_handle {
    int device = 0;
    int n_channels = 0;
    int max_chunk = 0;
    long long max_syms = 0;             // per-channel output stride of the staging buffers
    tdm_config cfg{};
    tdm_design design{};
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // host -> device staging of TDM_MEM_HOST calls, ahead of the kernels
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t ev_slice[16] = {};               // slice k of the capture has landed in d_iq
    int variant = 0;
    long long launches = 0;
    // device memory owned by the handle
    float* d_bank = nullptr;
    tdm_channel_state* d_states = nullptr;
    // staging for TDM_MEM_HOST callers (allocated lazily)
    float2* d_iq = nullptr;
    float2* d_syms = nullptr;
    uint8_t* d_dibits = nullptr;
    uint8_t* d_bits = nullptr;
    uint8_t* d_packed = nullptr;
    int* d_counts = nullptr;
    // tdm_process_long: the logical channel's carried state [0] + a freshly initialised state [1]; per-segment scratch
    tdm_channel_state* d_long_state = nullptr;   // [n_channels + 1]: carried state per logical channel, then one fresh state
    int long_channels = 0;                       // logical channels of the previous tdm_process_long_batch call
    tdm_channel_state* d_states2 = nullptr;
    uint8_t* d_seg_dibits = nullptr;
    uint8_t* d_seg_dibits2 = nullptr;
    long long seg_stride = 0, seg_stride2 = 0;
    int* d_seg_ints = nullptr;                   // counts[S], counts2[S], join[S], fixed[S], adopt[S], n_open, tail count, n_forced
    long long* d_offs = nullptr;                 // [rows] offsets, then [rows] per-channel totals
};

namespace {

void fill_params(const tdm_handle* h, tdm::DemodParams& p) {
    const tdm_design& d = h->design;
    std::memcpy(p.be_a, d.be_a, sizeof(p.be_a));
    std::memcpy(p.be_b, d.be_b, sizeof(p.be_b));
    std::memcpy(p.rrc, d.rrc, sizeof(p.rrc));
    p.agc_rate = d.agc_rate; p.agc_set_point = d.agc_set_point; p.agc_max_gain = d.agc_max_gain;
    p.fll_beta = d.fll_beta; p.fll_min_freq = d.fll_min_freq; p.fll_max_freq = d.fll_max_freq;
    p.tr_alpha = d.tr_alpha; p.tr_beta = d.tr_beta; p.tr_min_omega = d.tr_min_omega; p.tr_max_omega = d.tr_max_omega;
    p.costas_alpha = d.costas_alpha; p.costas_beta = d.costas_beta;
    p.costas_min_freq = d.costas_min_freq; p.costas_max_freq = d.costas_max_freq;
    p.fastamp_re_only = d.fastamp_re_only;
    p.packed = nullptr;
    p.packed_stride = 0;
    p.bank = h->d_bank;
    p.n_channels = h->n_channels;
    p.states = h->d_states;
    p.rows_per_channel = 1;
    p.sample_stride = 1;
    p.channel_stride = 0;
    p.debug_mask = 0;
}

void init_state(const tdm_design& d, tdm_channel_state& s) {
    std::memset(&s, 0, sizeof(s));
    s.agc_gain = d.agc_init_gain;       // FastAGC::_gain = initGain [A.3]
    s.fll_freq = d.fll_init_freq;       // pcl.init(..., initFreq, ...) fll.cpp:26
    s.tr_omega = d.tr_init_omega;       // pcl.init(..., _omega, ...) complex_fd.cpp:22
}

long long max_symbols_for(const tdm_design& d, long long count) {
    // Every symbol advances the read offset by floor(mu) with mu' = mu + omega + alpha*e,
    // |e| <= 1, omega >= tr_min_omega: K symbols consume >= K*(omega_min - |alpha|) - 1 samples.
    double step = (double)d.tr_min_omega - (double)(d.tr_alpha < 0 ? -d.tr_alpha : d.tr_alpha);
    if (step < 1.0) { step = 1.0; }
    long long k = (long long)((double)(count + 1) / step) + 2;
    return (k + 15) & ~15LL;
}

int upload_design(tdm_handle* h) {
    TDM_CUDA(cudaMemcpyAsync(h->d_bank, h->design.bank, sizeof(h->design.bank), cudaMemcpyHostToDevice, h->stream));
    TDM_CUDA(cudaStreamSynchronize(h->stream));
    return TDM_OK;
}

int ensure_staging(tdm_handle* h, uint32_t flags) {
    const size_t C = (size_t)h->n_channels;
    if (!h->d_iq) { TDM_CUDA(cudaMalloc(&h->d_iq, sizeof(float2) * C * (size_t)h->max_chunk)); }
    if (!h->d_counts) { TDM_CUDA(cudaMalloc(&h->d_counts, sizeof(int) * C)); }
    if ((flags & TDM_OUT_SYMBOLS) && !h->d_syms) { TDM_CUDA(cudaMalloc(&h->d_syms, sizeof(float2) * C * (size_t)h->max_syms)); }
    if ((flags & TDM_OUT_DIBITS) && !h->d_dibits) { TDM_CUDA(cudaMalloc(&h->d_dibits, C * (size_t)h->max_syms)); }
    if ((flags & TDM_OUT_BITS) && !h->d_bits) { TDM_CUDA(cudaMalloc(&h->d_bits, 2 * C * (size_t)h->max_syms)); }
    if ((flags & TDM_OUT_PACKED) && !h->d_packed) { TDM_CUDA(cudaMalloc(&h->d_packed, C * (size_t)(h->max_syms / 4))); }
    return TDM_OK;
}

}  // namespace

extern "C" {

const char* tdm_last_error(void) { return g_err; }
int tdm_abi_version(void) { return TDM_ABI_VERSION; }

int tdm_create(const tdm_config* cfg, int32_t n_channels, int32_t max_chunk, int32_t device, tdm_handle** out) {
    if (!out) { return fail(TDM_ERR_ARG, "tdm_create: out is null"); }
    *out = nullptr;
    if (n_channels <= 0 || max_chunk <= 0) { return fail(TDM_ERR_ARG, "tdm_create: n_channels and max_chunk must be > 0"); }
    tdm_config local;
    if (!cfg) { tdm_default_config(&local); cfg = &local; }
    tdm_design design;
    int rc = tdm_design_from_config(cfg, &design);
    if (rc != TDM_OK) { return fail(rc, "tdm_create: configuration not supported (rrc_tap_count must be 1..%d)", TDM_MAX_TAPS); }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        return fail(TDM_ERR_NO_DEVICE, "tdm_create: no CUDA device (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) { return fail(TDM_ERR_ARG, "tdm_create: device %d out of range (%d devices)", device, ndev); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { return fail(TDM_ERR_NO_DEVICE, "tdm_create: cannot query device %d", device); }
    if (prop.major != 10) {
        return fail(TDM_ERR_NO_DEVICE, "tdm_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    }

    tdm_handle* h = new (std::nothrow) tdm_handle();
    if (!h) { return fail(TDM_ERR_NOMEM, "tdm_create: out of host memory"); }
    h->device = device;
    h->n_channels = n_channels;
    h->max_chunk = max_chunk;
    h->cfg = *cfg;
    h->design = design;
    h->max_syms = max_symbols_for(design, max_chunk);
    DeviceGuard guard(device);
    auto cleanup = [&](int code) { tdm_destroy(h); return code; };
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { return cleanup(fail(TDM_ERR_CUDA, "cudaStreamCreate failed")); }
    h->stream = h->own_stream;
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { return cleanup(fail(TDM_ERR_CUDA, "cudaStreamCreate failed")); }
    for (auto& e : h->ev_slice) {
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { return cleanup(fail(TDM_ERR_CUDA, "cudaEventCreate failed")); }
    }
    if (cudaEventCreate(&h->ev_start) != cudaSuccess || cudaEventCreate(&h->ev_stop) != cudaSuccess) { return cleanup(fail(TDM_ERR_CUDA, "cudaEventCreate failed")); }
    if (cudaMalloc(&h->d_bank, sizeof(design.bank)) != cudaSuccess) { return cleanup(fail(TDM_ERR_NOMEM, "cudaMalloc(bank) failed")); }
    if (cudaMalloc(&h->d_states, sizeof(tdm_channel_state) * (size_t)n_channels) != cudaSuccess) { return cleanup(fail(TDM_ERR_NOMEM, "cudaMalloc(states) failed")); }
    rc = upload_design(h);
    if (rc != TDM_OK) { return cleanup(rc); }
    rc = tdm_reset_all(h);
    if (rc != TDM_OK) { return cleanup(rc); }
    *out = h;
    return TDM_OK;
}

int tdm_destroy(tdm_handle* h) {
    if (!h) { return TDM_OK; }
    DeviceGuard guard(h->device);
    if (h->own_stream) { cudaStreamSynchronize(h->own_stream); }
    cudaFree(h->d_bank); cudaFree(h->d_states); cudaFree(h->d_iq); cudaFree(h->d_syms);
    cudaFree(h->d_dibits); cudaFree(h->d_bits); cudaFree(h->d_packed); cudaFree(h->d_counts);
    cudaFree(h->d_long_state); cudaFree(h->d_states2); cudaFree(h->d_seg_dibits); cudaFree(h->d_seg_dibits2);
    cudaFree(h->d_seg_ints); cudaFree(h->d_offs);
    if (h->ev_start) { cudaEventDestroy(h->ev_start); }
    if (h->ev_stop) { cudaEventDestroy(h->ev_stop); }
    for (auto& e : h->ev_slice) { if (e) { cudaEventDestroy(e); } }
    if (h->copy_stream) { cudaStreamDestroy(h->copy_stream); }
    if (h->own_stream) { cudaStreamDestroy(h->own_stream); }
    delete h;
    return TDM_OK;
}

int tdm_set_stream(tdm_handle* h, void* cuda_stream) {
    if (!h) { return fail(TDM_ERR_ARG, "null handle"); }
    h->stream = (cuda_stream == TDM_OWN_STREAM) ? h->own_stream : (cudaStream_t)cuda_stream;
    return TDM_OK;
}

int64_t tdm_max_symbols(const tdm_handle* h, int64_t count) {
    if (!h || count < 0) { return TDM_ERR_ARG; }
    return max_symbols_for(h->design, count);
}

int tdm_process(tdm_handle* h, const float* iq, int64_t in_stride, int32_t count, float* syms, uint8_t* dibits,
                uint8_t* bits, int64_t out_stride, int32_t* out_counts, uint32_t out_flags, int32_t mem_kind) {
    if (out_flags & TDM_OUT_PACKED) { return fail(TDM_ERR_ARG, "tdm_process: TDM_OUT_PACKED needs tdm_process_io (it carries the packed buffer)"); }
    tdm_io io;
    std::memset(&io, 0, sizeof(io));
    io.iq = iq; io.in_stride = in_stride; io.count = count;
    io.syms = syms; io.dibits = dibits; io.bits = bits; io.out_stride = out_stride;
    io.out_counts = out_counts; io.out_flags = out_flags; io.mem_kind = mem_kind;
    return tdm_process_io(h, &io);
}

int tdm_process_io(tdm_handle* h, const tdm_io* io) {
    if (!h || !io) { return fail(TDM_ERR_ARG, "tdm_process: null handle / io"); }
    const float* iq = io->iq;
    const int64_t in_stride = io->in_stride, out_stride = io->out_stride, packed_stride = io->packed_stride;
    const int32_t count = io->count, mem_kind = io->mem_kind;
    float* syms = io->syms;
    uint8_t *dibits = io->dibits, *bits = io->bits, *packed = io->packed;
    int32_t* out_counts = io->out_counts;
    const uint32_t out_flags = io->out_flags;
    if (count < 0) { return fail(TDM_ERR_ARG, "tdm_process: negative count"); }
    if (!out_counts) { return fail(TDM_ERR_ARG, "tdm_process: out_counts is null"); }
    if (count > 0 && !iq) { return fail(TDM_ERR_ARG, "tdm_process: iq is null"); }
    const long long sstride = io->sample_stride ? (long long)io->sample_stride : 1;
    if (sstride == 1 && in_stride < count) { return fail(TDM_ERR_ARG, "tdm_process: in_stride < count"); }
    if (sstride != 1 && mem_kind != TDM_MEM_DEVICE) { return fail(TDM_ERR_ARG, "tdm_process: sample_stride needs TDM_MEM_DEVICE (host buffers are staged row by row)"); }
    if (sstride != 1 && in_stride < 1) { return fail(TDM_ERR_ARG, "tdm_process: in_stride < 1"); }
    if (out_flags & ~(TDM_OUT_SYMBOLS | TDM_OUT_DIBITS | TDM_OUT_BITS | TDM_OUT_PACKED)) { return fail(TDM_ERR_ARG, "tdm_process: unknown bits in out_flags"); }
    if (((out_flags & TDM_OUT_SYMBOLS) && !syms) || ((out_flags & TDM_OUT_DIBITS) && !dibits) || ((out_flags & TDM_OUT_BITS) && !bits) ||
        ((out_flags & TDM_OUT_PACKED) && !packed)) {
        return fail(TDM_ERR_ARG, "tdm_process: an output selected in out_flags has a null buffer");
    }
    const long long need = max_symbols_for(h->design, count);
    if ((out_flags & (TDM_OUT_SYMBOLS | TDM_OUT_DIBITS | TDM_OUT_BITS)) && out_stride < need) {
        return fail(TDM_ERR_ARG, "tdm_process: out_stride %lld < tdm_max_symbols(count) = %lld", (long long)out_stride, need);
    }
    if ((out_flags & TDM_OUT_PACKED) && packed_stride < need / 4) {
        return fail(TDM_ERR_ARG, "tdm_process: packed_stride %lld < tdm_max_symbols(count) / 4 = %lld", (long long)packed_stride, need / 4);
    }
    DeviceGuard guard(h->device);
    const size_t C = (size_t)h->n_channels;

    tdm::DemodParams p;
    fill_params(h, p);
    p.count = count;

    if (mem_kind == TDM_MEM_DEVICE) {
        p.iq = reinterpret_cast<const float2*>(iq);
        p.in_stride = in_stride;
        p.sample_stride = sstride;
        p.syms = (out_flags & TDM_OUT_SYMBOLS) ? reinterpret_cast<float2*>(syms) : nullptr;
        p.dibits = (out_flags & TDM_OUT_DIBITS) ? dibits : nullptr;
        p.bits = (out_flags & TDM_OUT_BITS) ? bits : nullptr;
        p.packed = (out_flags & TDM_OUT_PACKED) ? packed : nullptr;
        p.packed_stride = packed_stride;
        // only the packed output selected: rows are limited by what a call of `count` samples can emit
        p.out_stride = (out_flags & (TDM_OUT_SYMBOLS | TDM_OUT_DIBITS | TDM_OUT_BITS)) ? out_stride : need;
        p.out_counts = out_counts;
        p.accumulate = 0;
        TDM_CUDA(cudaEventRecord(h->ev_start, h->stream));
        const int n = tdm::launch_demod(p, h->variant, h->stream);
        if (n < 0) { return fail(TDM_ERR_CUDA, "demod kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
        h->launches += n;
        TDM_CUDA(cudaEventRecord(h->ev_stop, h->stream));
        return TDM_OK;
    }
    if (mem_kind != TDM_MEM_HOST) { return fail(TDM_ERR_ARG, "tdm_process: unknown mem_kind %d", mem_kind); }
    if (count > h->max_chunk) { return fail(TDM_ERR_ARG, "tdm_process: count %d > max_chunk %d given to tdm_create", count, h->max_chunk); }

    int rc = ensure_staging(h, out_flags);
    if (rc != TDM_OK) { return rc; }
    const long long dstride = h->max_syms;      // staging rows are max_syms wide (a multiple of 16)
    p.syms = (out_flags & TDM_OUT_SYMBOLS) ? h->d_syms : nullptr;
    p.dibits = (out_flags & TDM_OUT_DIBITS) ? h->d_dibits : nullptr;
    p.bits = (out_flags & TDM_OUT_BITS) ? h->d_bits : nullptr;
    p.packed = (out_flags & TDM_OUT_PACKED) ? h->d_packed : nullptr;
    p.packed_stride = dstride / 4;
    p.out_stride = dstride;
    p.out_counts = h->d_counts;
    p.in_stride = count;
    // The capture is cut into time slices: slice k+1 crosses PCIe (copy stream) while slice k is demodulated
    // (handle stream).  The chain is chunk invariant bit for bit -- every kernel carries its state in
    // tdm_channel_state -- so slicing changes nothing but the overlap; slices after the first append to the
    // output rows (DemodParams::accumulate).  A call is PCIe bound (8 B in per sample against 0.5 B out), so
    // hiding the kernel behind the copy is all there is to win.
    constexpr int kMaxSlices = (int)(sizeof(h->ev_slice) / sizeof(h->ev_slice[0]));
    int slice_len = count;
    if (count >= 16384) {
        slice_len = (count + 7) / 8;
        if (slice_len < 8192) { slice_len = 8192; }
        slice_len = (slice_len + 63) & ~63;                  // whole 8-sample blocks, aligned rows
    }
    const int n_slices = count > 0 ? (count + slice_len - 1) / slice_len : 1;
    if (n_slices > kMaxSlices) { return fail(TDM_ERR_ARG, "tdm_process: internal slice count"); }
    // the copy stream must not overtake work already queued on the handle's stream that still reads d_iq
    TDM_CUDA(cudaEventRecord(h->ev_slice[0], h->stream));
    TDM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_slice[0], 0));
    TDM_CUDA(cudaEventRecord(h->ev_start, h->stream));
    for (int k = 0; k < n_slices; ++k) {
        const int off = k * slice_len;
        const int len = (count - off < slice_len) ? count - off : slice_len;
        if (len > 0) {
            TDM_CUDA(cudaMemcpy2DAsync(h->d_iq + off, sizeof(float2) * (size_t)count, iq + 2 * (size_t)off, sizeof(float2) * (size_t)in_stride,
                                       sizeof(float2) * (size_t)len, C, cudaMemcpyHostToDevice, h->copy_stream));
        }
        TDM_CUDA(cudaEventRecord(h->ev_slice[k], h->copy_stream));
        TDM_CUDA(cudaStreamWaitEvent(h->stream, h->ev_slice[k], 0));
        p.iq = h->d_iq + off;
        p.count = len;
        p.accumulate = (k > 0) ? 1 : 0;
        const int n = tdm::launch_demod(p, h->variant, h->stream);
        if (n < 0) { return fail(TDM_ERR_CUDA, "demod kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
        h->launches += n;
    }
    TDM_CUDA(cudaEventRecord(h->ev_stop, h->stream));
    // Only the part of each row a call of `count` samples can fill is copied back.
    const size_t w = (size_t)need;
    if (out_flags & TDM_OUT_SYMBOLS) {
        TDM_CUDA(cudaMemcpy2DAsync(syms, sizeof(float2) * (size_t)out_stride, h->d_syms, sizeof(float2) * (size_t)dstride,
                                   sizeof(float2) * w, C, cudaMemcpyDeviceToHost, h->stream));
    }
    if (out_flags & TDM_OUT_DIBITS) {
        TDM_CUDA(cudaMemcpy2DAsync(dibits, (size_t)out_stride, h->d_dibits, (size_t)dstride, w, C, cudaMemcpyDeviceToHost, h->stream));
    }
    if (out_flags & TDM_OUT_BITS) {
        TDM_CUDA(cudaMemcpy2DAsync(bits, 2 * (size_t)out_stride, h->d_bits, 2 * (size_t)dstride, 2 * w, C, cudaMemcpyDeviceToHost, h->stream));
    }
    if (out_flags & TDM_OUT_PACKED) {
        TDM_CUDA(cudaMemcpy2DAsync(packed, (size_t)packed_stride, h->d_packed, (size_t)(dstride / 4), w / 4, C, cudaMemcpyDeviceToHost, h->stream));
    }
    TDM_CUDA(cudaMemcpyAsync(out_counts, h->d_counts, sizeof(int) * C, cudaMemcpyDeviceToHost, h->stream));
    TDM_CUDA(cudaStreamSynchronize(h->stream));
    return TDM_OK;
}

// SURVEY.md 8f rank 4.  See tdm_stitch.cu for the scheme and include/tdm_b200.h for the contract.
int tdm_process_long_batch(tdm_handle* h, const float* iq, int64_t in_stride, int64_t n_samples, int32_t n_channels, int32_t warmup,
                           uint8_t* dibits, int64_t out_stride, int64_t* out_counts, tdm_long_info* info, int32_t mem_kind) {
    if (!h || !info) { return fail(TDM_ERR_ARG, "tdm_process_long: null handle / info"); }
    std::memset(info, 0, sizeof(*info));
    if (n_channels < 1 || n_channels > h->n_channels) { return fail(TDM_ERR_ARG, "tdm_process_long: n_channels must be 1..%d (the handle's rows)", h->n_channels); }
    if (n_samples < 0 || in_stride < n_samples || out_stride < 0 || !out_counts || (n_samples > 0 && (!iq || !dibits))) { return fail(TDM_ERR_ARG, "tdm_process_long: bad buffers"); }
    if (warmup < 1024 || warmup > (1 << 24)) { return fail(TDM_ERR_ARG, "tdm_process_long: warmup must be 1024..2^24 samples"); }
    if (mem_kind != TDM_MEM_HOST && mem_kind != TDM_MEM_DEVICE) { return fail(TDM_ERR_ARG, "tdm_process_long: mem_kind"); }
    DeviceGuard guard(h->device);
    cudaStream_t st = h->stream;
    const int C = n_channels;
    if (n_samples == 0) {
        if (mem_kind == TDM_MEM_HOST) { for (int c = 0; c < C; ++c) { out_counts[c] = 0; } }
        else { TDM_CUDA(cudaMemsetAsync(out_counts, 0, sizeof(int64_t) * (size_t)C, st)); }
        return TDM_OK;
    }
    // dibits that must agree at a join.  128 would do to identify the place (2^-256 for a chance match), but a chain that
    // has only just locked still makes a stray decision error every few hundred symbols for a few thousand symbols
    // more (measured: 60 of 4096 segments at 1e9 samples with K = 128): demanding W/8 (up to 4096) error-free
    // symbols before the hand-over sends those segments to the sequential redo instead.
    int K = warmup / 8;
    K = K < 128 ? 128 : (K > 4096 ? 4096 : K);
    // S segments per channel: rows of L + W samples, segment s starts at sample s L.  A launch lasts as long as its
    // longest row ((L + W) samples at one recurrence's pace), so more, shorter segments pay as long as L stays well
    // above W: below 2 W the warm-up is more than a third of the work.
    int S = h->n_channels / C;
    const long long min_seg = 2LL * warmup;
    if ((n_samples - warmup) / S < min_seg) { S = (int)((n_samples - warmup) / min_seg); }
    if (S < 1) { S = 1; }
    const long long L = (S > 1) ? (((n_samples - warmup) / S) & ~7LL) : n_samples;
    const long long W = (S > 1) ? warmup : 0;
    if (L + W > 0x7fffffffLL) { return fail(TDM_ERR_ARG, "tdm_process_long: segments of %lld samples exceed the 32-bit count of a launch; use more rows", L + W); }
    const long long covered = (S > 1) ? S * L + W : n_samples;
    const long long tail = n_samples - covered;                 // < 9 S samples per channel, demodulated sequentially at the end
    // extension rounds: predecessors of segments that have not converged in time run on a little past their end and the
    // join is tried again further in (costs kExtSamples at one recurrence's pace instead of a whole segment's redo)
    constexpr int kExtRounds = 2;
    const long long X = (S > 1) ? ((L / 4 < 32768 ? L / 4 : 32768) & ~7LL) : 0;
    const long long need = max_symbols_for(h->design, L + W + kExtRounds * X);
    const int R = C * S;

    // scratch (sized for the handle's rows once)
    const size_t HR = (size_t)h->n_channels;
    if (!h->d_long_state) { TDM_CUDA(cudaMalloc(&h->d_long_state, (HR + 1) * sizeof(tdm_channel_state))); }
    if (h->long_channels != C) {                                // first use, after tdm_reset_all, or a different channel count: start from reset
        std::vector<tdm_channel_state> init((size_t)C + 1);
        for (auto& s : init) { init_state(h->design, s); }
        TDM_CUDA(cudaMemcpyAsync(h->d_long_state, init.data(), sizeof(tdm_channel_state) * init.size(), cudaMemcpyHostToDevice, st));
        TDM_CUDA(cudaStreamSynchronize(st));
        h->long_channels = C;
    }
    tdm_channel_state* d_carried = h->d_long_state;
    tdm_channel_state* d_fresh = h->d_long_state + C;
    if (!h->d_states2) { TDM_CUDA(cudaMalloc(&h->d_states2, sizeof(tdm_channel_state) * HR)); }
    if (!h->d_seg_ints) { TDM_CUDA(cudaMalloc(&h->d_seg_ints, sizeof(int) * (10 * HR + 4))); }
    if (!h->d_offs) { TDM_CUDA(cudaMalloc(&h->d_offs, sizeof(long long) * (2 * HR + 2))); }
    if (h->seg_stride < need) {
        cudaFree(h->d_seg_dibits); h->d_seg_dibits = nullptr; h->seg_stride = 0;
        TDM_CUDA(cudaMalloc(&h->d_seg_dibits, HR * (size_t)need));
        h->seg_stride = need;
    }
    int* d_counts = h->d_seg_ints;
    int* d_counts2 = d_counts + HR;
    int* d_join = d_counts2 + HR;
    int* d_fixed = d_join + HR;
    int* d_adopt = d_fixed + HR;
    int* d_agree = d_adopt + HR;
    int* d_mode = d_agree + HR;
    int* d_cut = d_mode + HR;
    int* d_tails = d_cut + HR;
    int* d_tailcount = d_tails + HR;                             // [C]
    int* d_nopen = d_tailcount + HR;
    int* d_nforced = d_nopen + 1;
    long long* d_offs = h->d_offs;
    long long* d_totals = h->d_offs + HR;

    const float2* d_iq = reinterpret_cast<const float2*>(iq);
    long long ch_stride = in_stride;
    float2* d_tmp_iq = nullptr;
    uint8_t* d_out = dibits;
    uint8_t* d_tmp_out = nullptr;
    long long* d_tmp_counts = nullptr;
    auto done = [&](int code) { cudaFree(d_tmp_iq); cudaFree(d_tmp_out); cudaFree(d_tmp_counts); return code; };
    if (mem_kind == TDM_MEM_HOST) {
        if (cudaMalloc(&d_tmp_iq, sizeof(float2) * (size_t)n_samples * (size_t)C) != cudaSuccess ||
            cudaMalloc(&d_tmp_out, (size_t)C * (size_t)(out_stride > 0 ? out_stride : 1)) != cudaSuccess) {
            return done(fail(TDM_ERR_NOMEM, "tdm_process_long: cannot stage %d x %lld samples on the device", C, (long long)n_samples));
        }
        if (cudaMemcpy2DAsync(d_tmp_iq, sizeof(float2) * (size_t)n_samples, iq, sizeof(float2) * (size_t)in_stride, sizeof(float2) * (size_t)n_samples,
                              (size_t)C, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            return done(fail(TDM_ERR_CUDA, "tdm_process_long: H2D copy failed"));
        }
        d_iq = d_tmp_iq; ch_stride = n_samples;
        d_out = d_tmp_out;
    }

    tdm::DemodParams p;
    fill_params(h, p);
    p.n_channels = R;
    p.syms = nullptr; p.bits = nullptr; p.accumulate = 0;
    p.rows_per_channel = S; p.channel_stride = ch_stride;
    // ---- pass 1: all segments at once, the first of every channel from its carried state, the others from reset state
    tdm::launch_long_init_states(h->d_states, d_carried, d_fresh, R, S, st);
    p.iq = d_iq; p.in_stride = (S > 1) ? L : ch_stride; p.count = (int)(L + W);
    p.dibits = h->d_seg_dibits; p.out_stride = h->seg_stride; p.out_counts = d_counts;
    int n = tdm::launch_demod(p, h->variant, st);
    if (n < 0) { return done(fail(TDM_ERR_CUDA, "demod kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()))); }
    h->launches += n;
    // ---- joins; segments that find none are redone as the continuation of their predecessor, one per failed run and pass
    const int mid = (int)(W / 2);
    int n_rerun = 0;
    if (S > 1) {
        if (cudaMemsetAsync(d_fixed, 0, sizeof(int) * (size_t)R, st) != cudaSuccess || cudaMemsetAsync(d_nforced, 0, sizeof(int), st) != cudaSuccess ||
            cudaMemsetAsync(d_cut, 0, sizeof(int) * (size_t)R, st) != cudaSuccess ||
            cudaMemcpyAsync(h->d_states2, h->d_states, sizeof(tdm_channel_state) * (size_t)R, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            return done(fail(TDM_ERR_CUDA, "tdm_process_long: %s", cudaGetErrorString(cudaGetLastError())));
        }
        tdm_channel_state* d_final = h->d_states2;                     // final loop state of the run whose stream each segment uses
        long long ext_total = 0;
        int ext_rounds = 0;
        for (int pass = 0;; ++pass) {
            tdm::launch_stitch_find(h->d_seg_dibits, h->seg_stride, d_counts, d_counts, R, S, K, mid - 2048, mid + 256, d_join, d_fixed, d_cut, 0, st);
            // A segment whose predecessor is not locked at the boundary (no signal there) is joined at the nominal place
            // right away; the pass limit only guards against pathological inputs.
            const bool give_up = pass >= 64;
            tdm::launch_stitch_plan(d_join, d_fixed, d_counts, R, S, d_adopt, d_nopen, d_nforced, mid, give_up ? 1 : 0, d_final, st);
            int n_open = 0;
            if (cudaMemcpyAsync(&n_open, d_nopen, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                return done(fail(TDM_ERR_CUDA, "tdm_process_long: %s", cudaGetErrorString(cudaGetLastError())));
            }
            if (n_open == 0) { break; }
            if (pass == 0) { n_rerun = n_open; }
            if (pass == 0 && X >= 4096 && R > 1) {
                // ---- extension rounds (only straight after pass 1: the rows' states are still their pass-1 finals).  Every
                // row but the very last runs on for X samples, appending to its own stream; an open segment then looks for
                // the extended predecessor's tail around symbol (W + ext) / 2 of its own stream.  The last row of the
                // buffer is left out (there is nothing behind it to read); other channel-final rows read into the next
                // channel, which is harmless: nobody joins them, and the carried states come from d_final.
                if (cudaMemcpyAsync(d_tails, d_counts, sizeof(int) * (size_t)R, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
                    return done(fail(TDM_ERR_CUDA, "tdm_process_long: copy failed"));
                }
                while (n_open > 0 && ext_rounds < kExtRounds) {
                    tdm::DemodParams e = p;
                    e.n_channels = R - 1;
                    e.iq = d_iq + (L + W) + ext_total; e.count = (int)X; e.accumulate = 1; e.out_counts = d_tails;
                    n = tdm::launch_demod(e, h->variant, st);
                    if (n < 0) { return done(fail(TDM_ERR_CUDA, "demod kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()))); }
                    h->launches += n;
                    ext_total += X; ++ext_rounds;
                    const int mid_e = (int)((W + ext_total) / 2);
                    int Ke = (int)((W + ext_total) / 8);                // the same rule as K, for the longer run-in
                    Ke = Ke < 128 ? 128 : (Ke > 4096 ? 4096 : Ke);
                    tdm::launch_stitch_find(h->d_seg_dibits, h->seg_stride, d_counts, d_tails, R, S, Ke, mid_e - 2048, mid_e + 256, d_join, d_fixed,
                                            d_cut, 1, st);
                    tdm::launch_stitch_plan(d_join, d_fixed, d_counts, R, S, d_adopt, d_nopen, d_nforced, mid, 0, d_final, st);
                    if (cudaMemcpyAsync(&n_open, d_nopen, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                        return done(fail(TDM_ERR_CUDA, "tdm_process_long: %s", cudaGetErrorString(cudaGetLastError())));
                    }
                }
                info->n_extended = n_rerun - n_open;
                if (n_open == 0) { break; }
            }
            const long long need2 = max_symbols_for(h->design, L);
            if (h->seg_stride2 < need2) {
                cudaFree(h->d_seg_dibits2); h->d_seg_dibits2 = nullptr; h->seg_stride2 = 0;
                if (cudaMalloc(&h->d_seg_dibits2, HR * (size_t)need2) != cudaSuccess) { return done(fail(TDM_ERR_NOMEM, "tdm_process_long: cudaMalloc failed")); }
                h->seg_stride2 = need2;
            }
            tdm::launch_long_shift_states(h->d_states, d_final, R, S, st);   // every segment continues its predecessor
            tdm::DemodParams q = p;
            q.iq = d_iq + W; q.count = (int)L;
            q.dibits = h->d_seg_dibits2; q.out_stride = h->seg_stride2; q.out_counts = d_counts2;
            n = tdm::launch_demod(q, h->variant, st);
            if (n < 0) { return done(fail(TDM_ERR_CUDA, "demod kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()))); }
            h->launches += n;
            tdm::launch_stitch_adopt(h->d_seg_dibits, h->seg_stride, h->d_seg_dibits2, h->seg_stride2, d_counts, d_counts2, d_join, d_fixed, d_adopt,
                                     d_agree, d_mode, d_nforced, mid, K < 512 ? K : 512, d_final, h->d_states, d_cut, R, S, need2, st);
        }
        // every logical channel continues from the run that produced its last segment's stream
        if (cudaMemcpyAsync(&info->n_forced, d_nforced, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
            return done(fail(TDM_ERR_CUDA, "tdm_process_long: copy failed"));
        }
        tdm::launch_long_last_states(d_carried, d_final, C, S, 0, st);
    } else {
        tdm::launch_long_last_states(d_carried, h->d_states, C, 1, 0, st);
    }
    if (S == 1 && cudaMemsetAsync(d_cut, 0, sizeof(int) * (size_t)R, st) != cudaSuccess) { return done(fail(TDM_ERR_CUDA, "tdm_process_long: memset failed")); }
    tdm::launch_stitch_scan(d_counts, d_cut, d_join, R, S, d_offs, d_totals, st);
    tdm::launch_stitch_copy(h->d_seg_dibits, h->seg_stride, d_counts, d_cut, d_join, d_offs, S, d_out, out_stride, R, need, st);
    // ---- the few samples the equal segments did not cover: sequentially, every channel from its last state (now in d_carried)
    if (tail > 0) {
        tdm::DemodParams q = p;
        q.n_channels = C; q.rows_per_channel = 1; q.channel_stride = 0;
        q.states = d_carried;
        q.iq = d_iq + covered; q.in_stride = ch_stride; q.count = (int)tail;
        q.dibits = h->d_seg_dibits; q.out_stride = h->seg_stride; q.out_counts = d_tailcount;
        n = tdm::launch_demod(q, h->variant, st);
        if (n < 0) { return done(fail(TDM_ERR_CUDA, "demod kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()))); }
        h->launches += n;
        tdm::launch_stitch_append(h->d_seg_dibits, h->seg_stride, d_tailcount, d_totals, d_out, out_stride, C, max_symbols_for(h->design, tail), st);
    }
    std::vector<long long> totals((size_t)C);
    if (cudaMemcpyAsync(totals.data(), d_totals, sizeof(long long) * (size_t)C, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        return done(fail(TDM_ERR_CUDA, "tdm_process_long: %s", cudaGetErrorString(cudaGetLastError())));
    }
    if (cudaGetLastError() != cudaSuccess) { return done(fail(TDM_ERR_CUDA, "tdm_process_long: kernel launch failed")); }
    long long sum = 0, worst = 0;
    for (long long t : totals) { sum += t; if (t > worst) { worst = t; } }
    info->n_dibits = sum; info->n_segments = S; info->n_rerun = n_rerun; info->segment_samples = (int32_t)L; info->warmup = (int32_t)W;
    static_assert(sizeof(long long) == sizeof(int64_t), "counts are copied as they are");
    if (mem_kind == TDM_MEM_HOST) {
        std::memcpy(out_counts, totals.data(), sizeof(int64_t) * (size_t)C);
    } else if (cudaMemcpyAsync(out_counts, d_totals, sizeof(int64_t) * (size_t)C, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        return done(fail(TDM_ERR_CUDA, "tdm_process_long: count copy failed"));
    }
    if (worst > out_stride) { return done(fail(TDM_ERR_ARG, "tdm_process_long: %lld dibits do not fit into out_stride %lld (rows truncated)", worst, (long long)out_stride)); }
    if (mem_kind == TDM_MEM_HOST) {
        if (cudaMemcpy2D(dibits, (size_t)out_stride, d_tmp_out, (size_t)out_stride, (size_t)worst, (size_t)C, cudaMemcpyDeviceToHost) != cudaSuccess) {
            return done(fail(TDM_ERR_CUDA, "tdm_process_long: D2H copy failed"));
        }
    } else if (cudaStreamSynchronize(st) != cudaSuccess) {
        return done(fail(TDM_ERR_CUDA, "tdm_process_long: %s", cudaGetErrorString(cudaGetLastError())));
    }
    return done(TDM_OK);
}

int tdm_process_long(tdm_handle* h, const float* iq, int64_t n_samples, int32_t warmup, uint8_t* dibits, int64_t dibits_cap,
                     tdm_long_info* info, int32_t mem_kind) {
    if (mem_kind == TDM_MEM_HOST) {
        int64_t count = 0;
        return tdm_process_long_batch(h, iq, n_samples, n_samples, 1, warmup, dibits, dibits_cap, &count, info, mem_kind);
    }
    if (!h) { return fail(TDM_ERR_ARG, "tdm_process_long: null handle"); }
    DeviceGuard guard(h->device);
    int64_t* d_count = nullptr;
    TDM_CUDA(cudaMalloc(&d_count, sizeof(int64_t)));
    const int rc = tdm_process_long_batch(h, iq, n_samples, n_samples, 1, warmup, dibits, dibits_cap, d_count, info, mem_kind);
    cudaFree(d_count);
    return rc;
}

int tdm_get_state(tdm_handle* h, tdm_channel_state* host_states, int32_t n_channels) {
    if (!h || !host_states || n_channels != h->n_channels) { return fail(TDM_ERR_ARG, "tdm_get_state: bad arguments"); }
    DeviceGuard guard(h->device);
    TDM_CUDA(cudaMemcpyAsync(host_states, h->d_states, sizeof(tdm_channel_state) * (size_t)n_channels, cudaMemcpyDeviceToHost, h->stream));
    TDM_CUDA(cudaStreamSynchronize(h->stream));
    return TDM_OK;
}

int tdm_set_state(tdm_handle* h, const tdm_channel_state* host_states, int32_t n_channels) {
    if (!h || !host_states || n_channels != h->n_channels) { return fail(TDM_ERR_ARG, "tdm_set_state: bad arguments"); }
    DeviceGuard guard(h->device);
    TDM_CUDA(cudaMemcpyAsync(h->d_states, host_states, sizeof(tdm_channel_state) * (size_t)n_channels, cudaMemcpyHostToDevice, h->stream));
    TDM_CUDA(cudaStreamSynchronize(h->stream));
    return TDM_OK;
}

int tdm_reset_all(tdm_handle* h) {
    if (!h) { return fail(TDM_ERR_ARG, "null handle"); }
    std::vector<tdm_channel_state> st((size_t)h->n_channels);
    for (auto& s : st) { init_state(h->design, s); }
    h->long_channels = 0;                        // the next tdm_process_long[_batch] starts every channel from reset state
    return tdm_set_state(h, st.data(), h->n_channels);
}

// PI4DQPSK::reset (pi4dqpsk.cpp:120-130) = fll.reset (phase 0, freq initFreq; fll.cpp:120-127),
// rrc.reset (FIR history cleared, [A.4]), agc.reset (gain = initGain), costas.reset (phase/freq
// back to their initial values, ph2 untouched), recov.reset (offset 0, mu 0, omega init;
// complex_fd.cpp:78-87).  One deliberate difference, forced by the shared delay line: the
// reference clears only the RRC FIR's history and keeps the two band-edge FIRs'; here the
// single line behind all three is cleared (DESIGN.md "reset").
int tdm_reset(tdm_handle* h) {
    if (!h) { return fail(TDM_ERR_ARG, "null handle"); }
    std::vector<tdm_channel_state> st((size_t)h->n_channels);
    int rc = tdm_get_state(h, st.data(), h->n_channels);
    if (rc != TDM_OK) { return rc; }
    for (auto& s : st) {
        s.fll_phase = 0; s.fll_freq = h->design.fll_init_freq;
        s.fll_quad = 0; s.fll_r = 0;            // the NCO's prepared reduction of phase 0 (tdm_math.cuh)
        std::memset(s.x_hist, 0, sizeof(s.x_hist));
        s.agc_gain = h->design.agc_init_gain;
        s.costas_phase = 0; s.costas_freq = 0;
        s.tr_offset = 0; s.tr_mu = 0; s.tr_omega = h->design.tr_init_omega;
    }
    return tdm_set_state(h, st.data(), h->n_channels);
}

int tdm_get_metrics(tdm_handle* h, tdm_metrics* m, int32_t n_channels) {
    if (!h || !m || n_channels != h->n_channels) { return fail(TDM_ERR_ARG, "tdm_get_metrics: bad arguments"); }
    std::vector<tdm_channel_state> st((size_t)n_channels);
    int rc = tdm_get_state(h, st.data(), n_channels);
    if (rc != TDM_OK) { return rc; }
    for (int c = 0; c < n_channels; ++c) {
        m[c].standarderr = st[c].standarderr; m[c].sync = st[c].sync;
        m[c].n_samples = st[c].n_samples; m[c].n_symbols = st[c].n_symbols;
    }
    return TDM_OK;
}

namespace {
// COMPLEX_FD::setOmega (complex_fd.cpp:31-42): offset = 0, pcl.phase = 0, pcl.freq = omega for every channel
int restart_timing(tdm_handle* h) {
    std::vector<tdm_channel_state> st((size_t)h->n_channels);
    int rc = tdm_get_state(h, st.data(), h->n_channels);
    if (rc != TDM_OK) { return rc; }
    for (auto& s : st) { s.tr_offset = 0; s.tr_mu = 0; s.tr_omega = h->design.tr_init_omega; }
    return tdm_set_state(h, st.data(), h->n_channels);
}
}  // namespace

int tdm_set_params(tdm_handle* h, const tdm_config* cfg, uint32_t what) {
    if (!h || !cfg) { return fail(TDM_ERR_ARG, "tdm_set_params: bad arguments"); }
    if (what & ~(TDM_SET_RATES | TDM_SET_RRC | TDM_SET_AGC_RATE | TDM_SET_COSTAS_BW | TDM_SET_FLL_BW | TDM_SET_TIMING_GAINS)) {
        return fail(TDM_ERR_ARG, "tdm_set_params: unknown bits in `what`");
    }
    tdm_design d;
    int rc = tdm_design_from_config(cfg, &d);
    if (rc != TDM_OK) { return fail(rc, "tdm_set_params: configuration not supported"); }
    DeviceGuard guard(h->device);
    TDM_CUDA(cudaStreamSynchronize(h->stream));
    tdm_design& cur = h->design;
    if (what & (TDM_SET_RATES | TDM_SET_RRC)) { std::memcpy(cur.rrc, d.rrc, sizeof(cur.rrc)); cur.ntaps = d.ntaps; }
    if (what & TDM_SET_AGC_RATE) { cur.agc_rate = d.agc_rate; }
    if (what & TDM_SET_COSTAS_BW) { cur.costas_alpha = d.costas_alpha; cur.costas_beta = d.costas_beta; }
    if (what & TDM_SET_FLL_BW) { cur.fll_beta = d.fll_beta; }
    if (what & TDM_SET_TIMING_GAINS) { cur.tr_alpha = d.tr_alpha; cur.tr_beta = d.tr_beta; cur.tr_min_omega = d.tr_min_omega; cur.tr_max_omega = d.tr_max_omega; }
    h->cfg = *cfg;
    if (what & TDM_SET_RATES) {
        cur.tr_init_omega = d.tr_init_omega; cur.tr_min_omega = d.tr_min_omega; cur.tr_max_omega = d.tr_max_omega;
        const long long ms = max_symbols_for(cur, h->max_chunk);
        if (ms > h->max_syms) {          // staging rows sized for the old rate are too short now: re-made lazily
            cudaFree(h->d_syms); cudaFree(h->d_dibits); cudaFree(h->d_bits); cudaFree(h->d_packed);
            h->d_syms = nullptr; h->d_dibits = nullptr; h->d_bits = nullptr; h->d_packed = nullptr;
        }
        h->max_syms = ms > h->max_syms ? ms : h->max_syms;
        return restart_timing(h);
    }
    return TDM_OK;
}

int tdm_set_config(tdm_handle* h, const tdm_config* cfg) {
    if (!h || !cfg) { return fail(TDM_ERR_ARG, "tdm_set_config: bad arguments"); }
    tdm_design d;
    int rc = tdm_design_from_config(cfg, &d);
    if (rc != TDM_OK) { return fail(rc, "tdm_set_config: configuration not supported"); }
    DeviceGuard guard(h->device);
    TDM_CUDA(cudaStreamSynchronize(h->stream));
    const bool rates_changed = cfg->samplerate != h->cfg.samplerate || cfg->symbolrate != h->cfg.symbolrate;
    h->cfg = *cfg;
    h->design = d;
    const long long ms = max_symbols_for(d, h->max_chunk);
    if (ms > h->max_syms) {
        // staging sized from the old design is too small now: drop it, it is re-made lazily
        cudaFree(h->d_syms); cudaFree(h->d_dibits); cudaFree(h->d_bits); cudaFree(h->d_packed);
        h->d_syms = nullptr; h->d_dibits = nullptr; h->d_bits = nullptr; h->d_packed = nullptr;
        h->max_syms = ms;
    }
    rc = upload_design(h);
    if (rc != TDM_OK) { return rc; }
    return rates_changed ? restart_timing(h) : TDM_OK;
}

int tdm_get_design(const tdm_handle* h, tdm_design* out) {
    if (!h || !out) { return fail(TDM_ERR_ARG, "tdm_get_design: bad arguments"); }
    *out = h->design;
    return TDM_OK;
}

int tdm_set_kernel_variant(tdm_handle* h, int32_t variant) {
    if (!h || variant < 0 || variant > tdm::kDemodVariants) { return fail(TDM_ERR_ARG, "tdm_set_kernel_variant: variant must be 0 (auto) .. %d", tdm::kDemodVariants); }
    h->variant = variant;
    return TDM_OK;
}

int tdm_last_kernel_ms(tdm_handle* h, float* ms) {
    if (!h || !ms) { return fail(TDM_ERR_ARG, "tdm_last_kernel_ms: bad arguments"); }
    DeviceGuard guard(h->device);
    TDM_CUDA(cudaEventSynchronize(h->ev_stop));
    TDM_CUDA(cudaEventElapsedTime(ms, h->ev_start, h->ev_stop));
    return TDM_OK;
}

int64_t tdm_launch_count(const tdm_handle* h) { return h ? h->launches : 0; }

int tdm_pack_dibits(tdm_handle* h, const uint8_t* dibits, int64_t in_stride, const int32_t* counts, uint8_t* packed,
                    int64_t out_stride) {
    if (!h || !dibits || !counts || !packed) { return fail(TDM_ERR_ARG, "tdm_pack_dibits: bad arguments"); }
    if (out_stride * 4 < in_stride) { return fail(TDM_ERR_ARG, "tdm_pack_dibits: out_stride too small"); }
    DeviceGuard guard(h->device);
    const int n = tdm::launch_pack_dibits(dibits, in_stride, counts, packed, out_stride, h->n_channels, in_stride, h->stream);
    if (n < 0) { return fail(TDM_ERR_CUDA, "pack kernel launch failed"); }
    h->launches += n;
    return TDM_OK;
}

int tdm_unpack_dibits(tdm_handle* h, const uint8_t* packed, int64_t in_stride, const int32_t* counts, int32_t n_rows,
                      uint8_t* dibits, int64_t dibit_stride, uint8_t* bits, int64_t bit_stride, int64_t max_symbols) {
    if (!h || !packed || !counts || (!dibits && !bits) || n_rows <= 0 || max_symbols < 0) { return fail(TDM_ERR_ARG, "tdm_unpack_dibits: bad arguments"); }
    if (in_stride * 4 < max_symbols || (dibits && dibit_stride < max_symbols) || (bits && bit_stride < 2 * max_symbols)) {
        return fail(TDM_ERR_ARG, "tdm_unpack_dibits: a row stride is smaller than max_symbols needs");
    }
    DeviceGuard guard(h->device);
    const int n = tdm::launch_unpack_dibits(packed, in_stride, counts, dibits, dibit_stride, bits, bit_stride, n_rows, max_symbols, h->stream);
    if (n < 0) { return fail(TDM_ERR_CUDA, "unpack kernel launch failed"); }
    h->launches += n;
    return TDM_OK;
}

int tdm_synth_capture(int32_t device, void* cuda_stream, const tdm_synth_params* p, int32_t n_channels, int64_t n_samples,
                      int64_t stride, int32_t first_channel, float* iq_dev, uint8_t* tx_dibits_dev, int64_t tx_stride) {
    if (!p || (!iq_dev && !tx_dibits_dev) || n_channels <= 0 || n_samples <= 0 || (iq_dev && stride < n_samples)) { return fail(TDM_ERR_ARG, "tdm_synth_capture: bad arguments"); }
    DeviceGuard guard(device);
    const int n = tdm::launch_synth(*p, n_channels, n_samples, stride, first_channel, reinterpret_cast<float2*>(iq_dev),
                                    tx_dibits_dev, tx_stride, (cudaStream_t)cuda_stream);
    if (n < 0) { return fail(TDM_ERR_CUDA, "synth kernel launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
    return TDM_OK;
}

}  // extern "C"
