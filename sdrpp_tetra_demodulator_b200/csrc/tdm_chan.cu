// tdm_chan.cu -- front-end channeliser: oversampled polyphase filterbank (include/tdm_chan_b200.h).
//
//   y_c[m] = sum_n h[n] x[t_m - n] e^{-j 2 pi c (t_m - n)/M},  t_m = (m+1) D - 1
//          = sum_p e^{+j 2 pi c (p - r_m)/M} v_m[p],   v_m[p] = sum_{q<T} h[p + q M] x[t_m - p - q M],   r_m = t_m mod M
// so one output instant = M branch sums (chan_residue_kernel), rotated by r_m, then an M-point inverse DFT (cuFFT, batched
// over instants).  The DFT leaves [instant][channel]: tdm_chan_process_instant_major hands that over as it is (the
// demodulator reads it in place, tdm_io.sample_stride), tdm_chan_process transposes it into channel-major rows.
//
// This stage is HBM-bound: per wideband sample 8 B are read once (every sample belongs to one residue slab, which stages
// it in shared memory once), the branch sums are written and read once (8 B * M / D each) and the channel samples
// written once (8 B * M / D): 8 + 3 * 8 * 36/25 = 42.6 B per wideband sample (the channel-major variant moves another
// 2 * 8 * M / D in its transposing pass).  Arithmetic: 2 T FMAs per branch sum (T = 16: 32) + 5 log2 M flops per channel
// sample -- against 390 FMAs per channel sample in the demodulator behind it.  Tensor cores: the DFT could be run as a
// [M x M] GEMM, but that is 8 M / (5 log2 M) = 30 .. 600 times the FFT's flops and would have to be split-TF32 to keep
// fp32's dynamic range next to a strong neighbour; an HBM-bound fp32 FFT is the right tool.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "tdm_chan_b200.h"
#include "tdm_b200.h"
#include "tdm_internal.h"

namespace {

// ---- cuFFT, loaded at run time (only cufftPlanMany / ExecC2C / SetStream / Destroy of the public C API)
typedef int cufftHandle;
typedef int cufftResult;
enum { CUFFT_C2C_ = 0x29, CUFFT_INVERSE_ = 1 };
struct Cufft {
    void* lib = nullptr;
    cufftResult (*PlanMany)(cufftHandle*, int, int*, int*, int, int, int*, int, int, int, int) = nullptr;
    cufftResult (*ExecC2C)(cufftHandle, float2*, float2*, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    bool ok = false;
};
Cufft& cufft() {
    static Cufft f;
    static std::once_flag once;
    std::call_once(once, [] {
        f.lib = dlopen("libcufft.so.11", RTLD_NOW | RTLD_GLOBAL);
        if (!f.lib) { f.lib = dlopen("libcufft.so", RTLD_NOW | RTLD_GLOBAL); }
        if (!f.lib) { return; }
#define TDM_SYM(field, name) f.field = reinterpret_cast<decltype(f.field)>(dlsym(f.lib, name))
        TDM_SYM(PlanMany, "cufftPlanMany");
        TDM_SYM(ExecC2C, "cufftExecC2C");
        TDM_SYM(SetStream, "cufftSetStream");
        TDM_SYM(Destroy, "cufftDestroy");
#undef TDM_SYM
        f.ok = f.PlanMany && f.ExecC2C && f.SetStream && f.Destroy;
    });
    return f;
}

double bessel_i0(double x) {
    double s = 1.0, t = 1.0;
    for (int k = 1; k < 60; ++k) { t *= (x / (2.0 * k)) * (x / (2.0 * k)); s += t; if (t < 1e-18 * s) { break; } }
    return s;
}

// Branch sums, organised by RESIDUE.  Branch sum v_m[p] reads the samples t_m - p - q M, q < T: all in the residue class
// rho = (t_m - p) mod M.  Seen from a residue class, the filterbank is an ordinary T-tap FIR over the class's own
// samples x_rho[j] = x[rho + j M] whose tap set changes with the instant (p_m = (t_m - rho) mod M), and consecutive
// instants reuse all but D/M of a sample on average.  So a CTA takes a slab of 32 residue classes and a chunk of
// instants, stages the slab's sample window in shared memory ONCE (rows of 32 consecutive wideband samples: 256-byte
// coalesced reads, each sample of the capture read by exactly one slab), and its threads (lane = residue class, warp =
// instant lane) form the branch sums from shared memory; the taps come through L1 (CTAs that run together share the
// slab, grid.x = chunk).  The first version had thread = branch and re-read every sample T M / D = 23 times from L2
// (1.39 ms per 16384 instants of 4608 channels, L2-bandwidth bound).  Sums run q = 0 .. T-1 as before: same bits.
constexpr int kSlab = 32;
constexpr int kChanRows = 192;               // rows of the staged window: 192 x 256 B = 48 KB
__device__ __forceinline__ long long floor_div(long long a, long long b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
// CS16: the capture arrives as interleaved int16 (re, im), what SDR hardware and baseband recordings deliver; sample value
// = s / 32768 (volk_16i_s32f_convert_32f(.., 32768.0f), src/dsp/osmotetra_dec.h:219 uses the same scale).  Half the
// bytes from HBM -- and across PCIe when the capture comes from the host.
__device__ __forceinline__ float2 cs16_to_float2(unsigned v) {
    return make_float2((float)(short)(v & 0xffffu) * (1.0f / 32768.0f), (float)(short)(v >> 16) * (1.0f / 32768.0f));
}
template <int T, bool CS16>
__global__ void __launch_bounds__(256) chan_residue_kernel(const void* __restrict__ in_raw, long long n_in, const float2* __restrict__ hist, long long n_hist,
                                                           const float* __restrict__ taps, int M, int D, long long n_out, int mi, int period,
                                                           long long t0_global /* global index of in[0] */, float2* __restrict__ u) {
    __shared__ float2 xs[kChanRows * kSlab];
    const float2* __restrict__ in = reinterpret_cast<const float2*>(in_raw);
    const unsigned* __restrict__ in16 = reinterpret_cast<const unsigned*>(in_raw);
    const int lane = threadIdx.x & 31, il = threadIdx.x >> 5;
    const int rho0 = blockIdx.y * kSlab, rho = rho0 + lane;
    const long long m0 = (long long)blockIdx.x * mi;
    const int cnt = (int)((n_out - m0) < mi ? (n_out - m0) : mi);
    // t_m - rho = base + (m - m0) D + (31 - lane) with base = t_{m0} - rho0 - 31; J = floor((t_m - rho) / M) = Jb + a / M, p = a % M,
    // a = rem + (m - m0) D + 31 - lane (32-bit: mi D < 2^30)
    const long long base = (m0 + 1) * D - 1 - rho0 - (kSlab - 1);
    const long long Jb = floor_div(base, M);
    const int rem = (int)(base - Jb * M);
    const long long Jlo = Jb - (T - 1);                                      // first staged row
    const int rows = (int)((rem + (long long)(cnt - 1) * D + (kSlab - 1)) / M) + T;      // <= kChanRows by the host's choice of mi
    if (rho < M) {
        // asynchronous copies (LDGSTS): every thread has all its rows in flight at once instead of a load -> store chain
        const uint32_t xs0 = (uint32_t)__cvta_generic_to_shared(xs) + 8u * (uint32_t)lane;
        for (int r = il; r < rows; r += 8) {
            const long long idx = (long long)rho + (Jlo + r) * M;            // local sample index: >= 0 new samples, < 0 carried history
            const float2* src = nullptr;
            if (idx >= 0) { if (idx < n_in) { if (CS16) { xs[r * kSlab + lane] = cs16_to_float2(__ldg(in16 + idx)); continue; } src = in + idx; } }
            else if (n_hist + idx >= 0) { src = hist + (n_hist + idx); }
            if (src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(xs0 + (uint32_t)(r * kSlab * 8)), "l"(src) : "memory"); }
            else { xs[r * kSlab + lane] = make_float2(0.f, 0.f); }
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (rho >= M) { return; }
    if (period > 0) {
        // period = M / gcd(D, M) instants later a class meets the SAME branch again (period D is a multiple of M; 36 for
        // every TETRA raster M = 36 g, D = 25 g): a warp takes the instants k = r (mod period) for a few r in turn, and the
        // T taps stay in registers for all of them -- shared memory then only carries the samples (128 B per branch sum)
        const int stepj = (int)(((long long)period * D) / M);
        for (int r = il; r < period && r < cnt; r += 8) {
            const unsigned a0 = (unsigned)rem + (unsigned)r * (unsigned)D + (unsigned)(kSlab - 1 - lane);
            int jr = (int)(a0 / (unsigned)M) + (T - 1);
            const int pbr = (int)(a0 % (unsigned)M);
            int pp = pbr - (int)((t0_global + (m0 + r + 1) * D - 1) % M);
            if (pp < 0) { pp += M; }
            float h[T];
#pragma unroll
            for (int q = 0; q < T; ++q) { h[q] = __ldg(taps + pbr + (long long)q * M); }
            for (int k = r; k < cnt; k += period) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < T; ++q) {
                    const float2 x = xs[(jr - q) * kSlab + lane];
                    acc.x = fmaf(h[q], x.x, acc.x);
                    acc.y = fmaf(h[q], x.y, acc.y);
                }
                u[(m0 + k) * M + pp] = acc;
                jr += stepj;
            }
        }
        return;
    }
    // general decimations: this thread's instants are m0 + il, m0 + il + 8, ...  (a, rotation) advance by 8 D per step, reduced without divisions
    const int stepq = (int)((8LL * D) / M), stepr = (int)((8LL * D) % M);
    const unsigned a0 = (unsigned)rem + (unsigned)il * (unsigned)D + (unsigned)(kSlab - 1 - lane);
    int jr = (int)(a0 / (unsigned)M) + (T - 1);                               // row of the newest sample (q = 0)
    int pbr = (int)(a0 % (unsigned)M);                                        // branch
    int rot = (int)((t0_global + (m0 + il + 1) * D - 1) % M);                 // r_m
    for (int k = il; k < cnt; k += 8) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < T; ++q) {
            const float h = __ldg(taps + pbr + (long long)q * M);
            const float2 x = xs[(jr - q) * kSlab + lane];
            acc.x = fmaf(h, x.x, acc.x);
            acc.y = fmaf(h, x.y, acc.y);
        }
        int pp = pbr - rot;
        if (pp < 0) { pp += M; }
        u[(m0 + k) * M + pp] = acc;
        pbr += stepr; jr += stepq;
        if (pbr >= M) { pbr -= M; ++jr; }
        rot += stepr;
        if (rot >= M) { rot -= M; }
    }
}

// [n_out][M] (what the batched DFT leaves, instant-major) -> [M][out_stride] (channel-major rows the demodulator reads in place)
__global__ void __launch_bounds__(256) chan_transpose_kernel(const float2* __restrict__ v, float2* __restrict__ out, int M, long long n_out, long long out_stride) {
    __shared__ float2 tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long m0 = (long long)blockIdx.x * 32;
    for (int k0 = blockIdx.y * 32; k0 < M; k0 += gridDim.y * 32) {
#pragma unroll
        for (int i = ty; i < 32; i += 8) {
            const long long m = m0 + i;
            const int k = k0 + tx;
            if (m < n_out && k < M) { tile[i][tx] = v[m * M + k]; }
        }
        __syncthreads();
#pragma unroll
        for (int i = ty; i < 32; i += 8) {
            const int k = k0 + i;
            const long long m = m0 + tx;
            if (m < n_out && k < M) { out[(long long)k * out_stride + m] = tile[tx][i]; }
        }
        __syncthreads();
    }
}

// the last n_hist samples of [hist | in] become the new history
template <bool CS16>
__global__ void chan_history_kernel(const void* __restrict__ in_raw, long long n_in, const float2* __restrict__ hist_old, float2* __restrict__ hist_new, long long n_hist) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_hist; i += (long long)gridDim.x * blockDim.x) {
        const long long src = n_in - n_hist + i;              // index into `in`; negative: old history
        if (src < 0) { hist_new[i] = hist_old[n_hist + src]; }
        else if (CS16) { hist_new[i] = cs16_to_float2(reinterpret_cast<const unsigned*>(in_raw)[src]); }
        else { hist_new[i] = reinterpret_cast<const float2*>(in_raw)[src]; }
    }
}

}  // namespace

struct tdm_chan {
    tdm_chan_config cfg{};
    int device = 0;
    std::vector<float> taps;
    float* d_taps = nullptr;
    float2* d_hist[2] = { nullptr, nullptr };
    int cur = 0;
    long long n_hist = 0;
    long long t_global = 0;          // wideband samples consumed so far
    float2* d_u = nullptr;
    long long u_cap = 0;             // instants the branch-sum buffer holds
    cufftHandle plan = 0;
    long long plan_batch = 0, plan_stride = 0;      // plan_stride: distance between consecutive instants of the DFT's output (M: in place)
    cudaEvent_t ev[3] = { nullptr, nullptr, nullptr };
};

extern "C" {

int tdm_chan_default_config(int32_t g, tdm_chan_config* cfg) {
    if (!cfg || g < 1) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_default_config: bad arguments"); }
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->n_channels = 36 * g;            // 25 kHz raster ...
    cfg->decimation = 25 * g;            // ... at 36 kS/s per channel (VFO_SAMPLERATE, src/main.cpp:35)
    cfg->taps_per_branch = 16;
    cfg->passband = 0.55;                // flat to +-13.75 kHz: a TETRA carrier occupies +-12.15 kHz (18 ksym/s, roll-off 0.35)
    cfg->stopband = 0.89;                // 22.25 kHz: what folds onto the carrier's +-12.15 kHz at 36 kS/s starts at 36 - 12.15 = 23.85 kHz
    cfg->stop_atten_db = 70.0;
    return TDM_OK;
}

int tdm_chan_design(const tdm_chan_config* cfg, float* taps) {
    if (!cfg || !taps) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_design: bad arguments"); }
    const int M = cfg->n_channels, T = cfg->taps_per_branch;
    if (M < 2 || T < 4 || T > 32 || cfg->decimation < 1 || cfg->decimation > M || !(cfg->passband > 0) || !(cfg->stopband > cfg->passband)) {
        return tdm_internal_fail(TDM_ERR_UNSUPPORTED, "tdm_chan_design: need M >= 2, 4 <= T <= 32, 1 <= D <= M, 0 < passband < stopband");
    }
    const long long L = (long long)T * M;
    const double A = cfg->stop_atten_db;
    const double beta = A > 50 ? 0.1102 * (A - 8.7) : (A > 21 ? 0.5842 * std::pow(A - 21, 0.4) + 0.07886 * (A - 21) : 0.0);
    const double fc = 0.5 * (cfg->passband + cfg->stopband) / M;          // cut-off in cycles per wideband sample
    const double pi = 3.14159265358979323846, mid = 0.5 * (double)(L - 1);
    double sum = 0.0;
    std::vector<double> h((size_t)L);
    for (long long n = 0; n < L; ++n) {
        const double t = (double)n - mid;
        const double s = t == 0.0 ? 2.0 * fc : std::sin(2.0 * pi * fc * t) / (pi * t);
        const double a = t / (mid + 0.5);
        const double w = bessel_i0(beta * std::sqrt(std::fmax(0.0, 1.0 - a * a))) / bessel_i0(beta);
        h[(size_t)n] = s * w;
        sum += h[(size_t)n];
    }
    for (long long n = 0; n < L; ++n) { taps[n] = (float)(h[(size_t)n] / sum); }
    return TDM_OK;
}

int tdm_chan_create(const tdm_chan_config* cfg, int32_t device, tdm_chan** out) {
    if (!out) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_create: out is null"); }
    *out = nullptr;
    tdm_chan_config local;
    if (!cfg) { tdm_chan_default_config(4, &local); cfg = &local; }
    const int M = cfg->n_channels, T = cfg->taps_per_branch;
    std::vector<float> taps;
    if (M >= 2 && T >= 4 && T <= 32) { taps.resize((size_t)T * (size_t)M); }
    int rc = tdm_chan_design(cfg, taps.data());
    if (rc != TDM_OK) { return rc; }
    if (T != 8 && T != 12 && T != 16 && T != 24 && T != 32) { return tdm_internal_fail(TDM_ERR_UNSUPPORTED, "tdm_chan_create: taps_per_branch must be 8, 12, 16, 24 or 32"); }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { return tdm_internal_fail(TDM_ERR_NO_DEVICE, "tdm_chan_create: no CUDA device (this library has no CPU fallback)"); }
    if (device < 0 || device >= ndev) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_create: device out of range"); }
    if (!cufft().ok) { return tdm_internal_fail(TDM_ERR_UNSUPPORTED, "tdm_chan_create: libcufft.so.11 could not be loaded"); }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    tdm_chan* c = new (std::nothrow) tdm_chan();
    if (!c) { return tdm_internal_fail(TDM_ERR_NOMEM, "tdm_chan_create: out of host memory"); }
    c->cfg = *cfg; c->device = device; c->taps = taps;
    c->n_hist = (long long)T * M;
    bool ok = cudaMalloc(&c->d_taps, sizeof(float) * taps.size()) == cudaSuccess &&
              cudaMalloc(&c->d_hist[0], sizeof(float2) * (size_t)c->n_hist) == cudaSuccess &&
              cudaMalloc(&c->d_hist[1], sizeof(float2) * (size_t)c->n_hist) == cudaSuccess &&
              cudaMemcpy(c->d_taps, taps.data(), sizeof(float) * taps.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemset(c->d_hist[0], 0, sizeof(float2) * (size_t)c->n_hist) == cudaSuccess;
    for (auto& e : c->ev) { ok = ok && cudaEventCreate(&e) == cudaSuccess; }
    if (prev >= 0 && prev != device) { cudaSetDevice(prev); }
    if (!ok) { tdm_chan_destroy(c); return tdm_internal_fail(TDM_ERR_NOMEM, "tdm_chan_create: device allocation failed"); }
    *out = c;
    return TDM_OK;
}

int tdm_chan_destroy(tdm_chan* c) {
    if (!c) { return TDM_OK; }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    if (c->plan && cufft().ok) { cufft().Destroy(c->plan); }
    cudaFree(c->d_taps); cudaFree(c->d_hist[0]); cudaFree(c->d_hist[1]); cudaFree(c->d_u);
    for (auto& e : c->ev) { if (e) { cudaEventDestroy(e); } }
    if (prev >= 0 && prev != c->device) { cudaSetDevice(prev); }
    delete c;
    return TDM_OK;
}

int tdm_chan_reset(tdm_chan* c) {
    if (!c) { return tdm_internal_fail(TDM_ERR_ARG, "null handle"); }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    const bool ok = cudaMemset(c->d_hist[c->cur], 0, sizeof(float2) * (size_t)c->n_hist) == cudaSuccess;
    if (prev >= 0 && prev != c->device) { cudaSetDevice(prev); }
    c->t_global = 0;
    return ok ? TDM_OK : tdm_internal_fail(TDM_ERR_CUDA, "tdm_chan_reset: memset failed");
}

static int chan_process(tdm_chan* c, const void* wide, bool cs16, int64_t n_wide, float* out, int64_t out_stride, bool instant_major, void* cuda_stream);
int tdm_chan_process(tdm_chan* c, const float* wide, int64_t n_wide, float* out, int64_t out_stride, void* cuda_stream) {
    return chan_process(c, wide, false, n_wide, out, out_stride, false, cuda_stream);
}
int tdm_chan_process_instant_major(tdm_chan* c, const float* wide, int64_t n_wide, float* out, int64_t row_pitch, void* cuda_stream) {
    return chan_process(c, wide, false, n_wide, out, row_pitch, true, cuda_stream);
}
int tdm_chan_process_ex(tdm_chan* c, const void* wide, int32_t in_format, int64_t n_wide, float* out, int64_t pitch, int32_t out_layout, void* cuda_stream) {
    if ((in_format != TDM_CHAN_IN_CF32 && in_format != TDM_CHAN_IN_CS16) || (out_layout != TDM_CHAN_OUT_CHANNEL_MAJOR && out_layout != TDM_CHAN_OUT_INSTANT_MAJOR)) {
        return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_process_ex: unknown in_format / out_layout");
    }
    return chan_process(c, wide, in_format == TDM_CHAN_IN_CS16, n_wide, out, pitch, out_layout == TDM_CHAN_OUT_INSTANT_MAJOR, cuda_stream);
}
static int chan_process(tdm_chan* c, const void* wide, bool cs16, int64_t n_wide, float* out, int64_t out_stride, bool instant_major, void* cuda_stream) {
    if (!c || n_wide < 0 || (n_wide > 0 && (!wide || !out))) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_process: bad arguments"); }
    const int M = c->cfg.n_channels, D = c->cfg.decimation, T = c->cfg.taps_per_branch;
    if (n_wide % D) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_process: n_wide must be a multiple of the decimation %d", D); }
    const long long n_out = n_wide / D;
    if (n_out == 0) { return TDM_OK; }
    if (!instant_major && out_stride < n_out) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_process: out_stride < n_wide / D"); }
    if (instant_major && (out_stride < M || out_stride > 0x7fffffffLL)) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_process_instant_major: row_pitch < M"); }
    if (n_out > 0x7fffffffLL) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_process: too many output instants for one call"); }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    auto leave = [&](int code) { if (prev >= 0 && prev != c->device) { cudaSetDevice(prev); } return code; };
    if (c->u_cap < n_out) {
        cudaFree(c->d_u); c->d_u = nullptr; c->u_cap = 0;
        if (cudaMalloc(&c->d_u, sizeof(float2) * (size_t)n_out * (size_t)M) != cudaSuccess) { return leave(tdm_internal_fail(TDM_ERR_NOMEM, "tdm_chan_process: cannot allocate the branch-sum buffer")); }
        c->u_cap = n_out;
    }
    const long long dft_dist = instant_major ? out_stride : M;
    if (!c->plan || c->plan_batch != n_out || c->plan_stride != dft_dist) {
        if (c->plan) { cufft().Destroy(c->plan); c->plan = 0; }
        int n[1] = { M }, inembed[1] = { M }, onembed[1] = { M };
        // instant m at u + m M, unit stride, in place; chan_transpose_kernel then writes channel k of instant m at out + k * out_stride + m
        // (cuFFT can write that layout itself -- ostride = out_stride, odist = 1 -- but took 0.89 ms for 16384 x 4608 that way)
        if (cufft().PlanMany(&c->plan, 1, n, inembed, 1, M, onembed, 1, (int)dft_dist, CUFFT_C2C_, (int)n_out) != 0) {
            c->plan = 0;
            return leave(tdm_internal_fail(TDM_ERR_CUDA, "tdm_chan_process: cufftPlanMany failed (M = %d, batch = %lld)", M, n_out));
        }
        c->plan_batch = n_out; c->plan_stride = dft_dist;
    }
    const void* in = wide;
    // instants per CTA: as many as the 192-row window holds (rows = ceil(((mi - 1) D + 31) / M) + T)
    long long mi = ((long long)(kChanRows - T - 1) * M - (kSlab - 1)) / D;
    mi = mi > 1024 ? 1024 : (mi < 8 ? 8 : (mi & ~7LL));
    const long long n_chunks = (n_out + mi - 1) / mi;
    dim3 grid((unsigned)n_chunks, (unsigned)((M + kSlab - 1) / kSlab));
    int gcd_dm = D, tmp = M;
    while (tmp) { const int t = gcd_dm % tmp; gcd_dm = tmp; tmp = t; }
    const int period = (M / gcd_dm <= 64 && M / gcd_dm <= mi / 2) ? M / gcd_dm : 0;     // short branch period: taps-in-registers path
    cudaEventRecord(c->ev[0], st);
#define TDM_CHAN_LAUNCH(TT) if (cs16) { chan_residue_kernel<TT, true><<<grid, 256, 0, st>>>(in, n_wide, c->d_hist[c->cur], c->n_hist, c->d_taps, M, D, n_out, (int)mi, period, c->t_global, c->d_u); } else chan_residue_kernel<TT, false><<<grid, 256, 0, st>>>(in, n_wide, c->d_hist[c->cur], c->n_hist, c->d_taps, M, D, n_out, (int)mi, period, c->t_global, c->d_u)
    switch (T) {
        case 8: TDM_CHAN_LAUNCH(8); break;
        case 12: TDM_CHAN_LAUNCH(12); break;
        case 16: TDM_CHAN_LAUNCH(16); break;
        case 24: TDM_CHAN_LAUNCH(24); break;
        default: TDM_CHAN_LAUNCH(32); break;
    }
#undef TDM_CHAN_LAUNCH
    cudaEventRecord(c->ev[1], st);
    if (cufft().SetStream(c->plan, st) != 0 || cufft().ExecC2C(c->plan, c->d_u, instant_major ? reinterpret_cast<float2*>(out) : c->d_u, CUFFT_INVERSE_) != 0) {
        return leave(tdm_internal_fail(TDM_ERR_CUDA, "tdm_chan_process: cuFFT execution failed"));
    }
    if (!instant_major) {
        const long long tiles_m = (n_out + 31) / 32;
        const int tiles_k = (M + 31) / 32;
        chan_transpose_kernel<<<dim3((unsigned)tiles_m, (unsigned)(tiles_k < 65535 ? tiles_k : 65535)), 256, 0, st>>>(c->d_u, reinterpret_cast<float2*>(out), M, n_out, out_stride);
    }
    cudaEventRecord(c->ev[2], st);
    if (cs16) { chan_history_kernel<true><<<256, 256, 0, st>>>(in, n_wide, c->d_hist[c->cur], c->d_hist[c->cur ^ 1], c->n_hist); }
    else { chan_history_kernel<false><<<256, 256, 0, st>>>(in, n_wide, c->d_hist[c->cur], c->d_hist[c->cur ^ 1], c->n_hist); }
    c->cur ^= 1;
    c->t_global += n_wide;
    if (cudaGetLastError() != cudaSuccess) { return leave(tdm_internal_fail(TDM_ERR_CUDA, "tdm_chan_process: kernel launch failed")); }
    return leave(TDM_OK);
}

int tdm_chan_last_kernel_ms(tdm_chan* c, float* polyphase_ms, float* dft_ms) {
    if (!c || !polyphase_ms || !dft_ms) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_chan_last_kernel_ms: bad arguments"); }
    if (cudaEventSynchronize(c->ev[2]) != cudaSuccess || cudaEventElapsedTime(polyphase_ms, c->ev[0], c->ev[1]) != cudaSuccess ||
        cudaEventElapsedTime(dft_ms, c->ev[1], c->ev[2]) != cudaSuccess) { return tdm_internal_fail(TDM_ERR_CUDA, "tdm_chan_last_kernel_ms: no call timed yet"); }
    return TDM_OK;
}

}  // extern "C"
