// tdm_kernels.cu -- the fused pi/4-DQPSK demodulation kernel, thread-per-channel mapping, and the
// dispatcher that picks a mapping for a launch.
//
// One launch runs the WHOLE reference chain for `count` samples of every channel:
//   FastAGC -> band-edge FLL -> RRC matched filter -> ML timing recovery -> pi/4 Costas
//   -> slicer + differential decoder -> (optional) bit unpack / 4-per-byte packing
// i.e. dsp::demod::PI4DQPSK::process (src/dsp/pi4dqpsk.cpp:132-140), then
// DQPSKSymbolExtractor::process (src/dsp/dqpsk_sym_extr.cpp:4-55) and
// BitUnpacker::process (src/dsp/bit_unpacker.cpp:4-10), with the per-channel state
// the reference keeps in class members carried in tdm_channel_state.
//
// Mapping "tpc<T>" (this file): one thread per channel, time processed in blocks of T
// samples.  The chain is a strict recurrence at sample rate (AGC gain, FLL phase) and
// at symbol rate (timing, Costas), so time cannot be split; what CAN be hoisted out of
// the recurrence is almost all of the FIR work:
//   * lbe/hbe/RRC all filter the same FLL-output sequence x (fll.cpp:141-142 ->
//     pi4dqpsk.cpp:135-136), so ONE 64-deep delay line (shared memory, [entry][lane] so
//     a warp's accesses are conflict-free) feeds all three;
//   * the band-edge taps are exact conjugates (fll.cpp:89-93): hbe = P + jQ, lbe = P - jQ
//     with P = sum a_k x, Q = sum b_k x  ->  6 real chains per output instead of 10;
//   * for a block of T outputs the contributions of the 64 samples older than the block
//     ("old part") do not depend on the block's own feedback: they are accumulated first,
//     each history sample loaded once from shared memory and used for up to T outputs
//     x 6 chains, taps as constant-bank FFMA operands.  Only the last <= T terms of each
//     chain sit inside the serial loop.
// Every chain still adds its terms in ascending tap order with one fma per term, which
// is the canonical order the CPU checker follows -- the blocking changes WHEN a term is
// added, never the order within a chain.
//
// This mapping is the simple one: every stage in one thread, in program order.  It is the
// in-library cross-check of the warp-specialised pipeline (tdm_ws.cu), which is what `auto`
// launches.
#include "tdm_chain.cuh"

namespace tdm {

namespace {

template <int T>
struct TpcLayout {
    static_assert(kHist % T == 0, "block length must divide the history length");
    static constexpr int kSlots = kHist / T + 1;         // ring of (64/T + 1) blocks of T
    static constexpr int kXEntries = kSlots * T;
    static constexpr int kREntries = (T + kITaps - 1 <= 16) ? 16 : 32;   // RRC-output ring (power of 2)
    static constexpr int kWarpFloat2 = (kXEntries + kREntries) * 32;
    static constexpr size_t kWarpBytes = sizeof(float2) * kWarpFloat2;
};

// Instruction-cache discipline: the first version of this kernel unrolled everything
// (53 KB of SASS per block iteration) and ncu showed `stall_no_instruction` as the top
// stall at IPC 0.36.  The hot loop is therefore kept ROLLED and small (about 10 KB):
//   * old part: a loop over the 64/T history slots; each iteration loads T history
//     samples and the 2T-1 taps per filter they meet (uniform constant-bank loads from
//     the T-1-zero-padded tap tables, tpad[f][s*T + j-i+T-1]) and does T*T*6 FMAs;
//   * serial part: a loop over the T samples; the T in-flight chains sit in a register
//     shift-register, so position q always meets tap 64-q (an immediate constant-bank
//     operand) and the loop body does not depend on the sample index;
//   * symbol part: timing -> Costas -> slicer, called from a while loop.
template <int T, bool RE_ONLY>
__global__ void __launch_bounds__(128) demod_tpc_kernel(const __grid_constant__ DemodParams p) {
    using L = TpcLayout<T>;
    constexpr int S = L::kSlots;
    constexpr int RE = L::kREntries;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* bank4 = reinterpret_cast<float4*>(smem_raw);                         // [128 * 2] float4 (plain table)
    float2* warp_base = reinterpret_cast<float2*>(smem_raw + sizeof(float) * kIPhases * kITaps) +
                        (size_t)(threadIdx.x >> 5) * L::kWarpFloat2;
    float2* xs = warp_base;                     // [kXEntries][32]
    float2* rs = warp_base + L::kXEntries * 32; // [RE][32]
    const int lane = threadIdx.x & 31;

    {
        const float4* __restrict__ b4 = reinterpret_cast<const float4*>(p.bank);
        for (int i = threadIdx.x; i < kIPhases * 2; i += blockDim.x) { bank4[i] = __ldg(b4 + i); }
    }

    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = ch < p.n_channels;
    if (!active) { ch = p.n_channels - 1; }     // compute a duplicate, store nothing
    tdm_channel_state* __restrict__ sp = p.states + ch;

    // ---- load carried state
    float g = sp->agc_gain;
    FllState fs = fll_load(sp);
    float mu = sp->tr_mu, om = sp->tr_omega;
    int offset = sp->tr_offset;
    float cph = sp->costas_phase, cfr = sp->costas_freq, ph2 = sp->costas_ph2;
    SlicerState sl;
    float err_blocks[TDM_SYNC_BLOCKS];
    slicer_load(p, sp, ch, sl, err_blocks);
    // delay line: linear index q (0..63 = carried history, 64+n = new sample n) lives in ring
    // slot (1 + q/T) mod S, position q%T;  RRC ring: linear q' (0..6 history, 7+n new) at q' & (RE-1)
    {
        const float2* xh = reinterpret_cast<const float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) { xs[((1 + m / T) * T + (m % T)) * 32 + lane] = xh[m]; }
        const float2* rh = reinterpret_cast<const float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { rs[j * 32 + lane] = rh[j]; }
    }
    __syncthreads();   // bank visible

    const float2* __restrict__ in = row_input(p, ch);
    const long long ss = p.sample_stride;
    const int count = p.count;
    const int nblk = (count + T - 1) / T;

    float2 cur[T], nxt[T];
#pragma unroll
    for (int i = 0; i < T; ++i) { cur[i] = (i < count) ? __ldg(in + i * ss) : make_float2(0.f, 0.f); }
    const SymConsts kc = load_sym_consts(p);
    const LoopConsts lc = load_loop_consts(p);

    int slot = 0;
#pragma unroll 1
    for (int blk = 0; blk < nblk; ++blk) {
        const int n0 = blk * T;
        const int valid = min(T, count - n0);
        // prefetch the next block's input while this one computes
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const int n = n0 + T + i;
            nxt[i] = (n < count) ? __ldg(in + n * ss) : make_float2(0.f, 0.f);
        }

        // ---- old part: history terms of all T outputs (independent of this block's feedback).
        // acc[i][*] is the chain of output i; history sample m = s*T + j meets it at tap m - i, which
        // is entry j - i + T - 1 of the padded table row starting at s*T (negative taps read zeros).
        float acc[T][6];
#pragma unroll
        for (int i = 0; i < T; ++i) {
#pragma unroll
            for (int c = 0; c < 6; ++c) { acc[i][c] = 0.f; }
        }
        {
            int hs = slot + 1;
#pragma unroll 1
            for (int s = 0; s < kHist / T; ++s) {
                if (hs >= S) { hs -= S; }
                float ta[2 * T - 1], tb[2 * T - 1], tr[2 * T - 1];
#pragma unroll
                for (int c = 0; c < 2 * T - 1; ++c) {
                    ta[c] = p.tpad[0][s * T + c];
                    tb[c] = p.tpad[1][s * T + c];
                    tr[c] = p.tpad[2][s * T + c];
                }
#pragma unroll
                for (int j = 0; j < T; ++j) {
                    const float2 h = xs[(hs * T + j) * 32 + lane];
#pragma unroll
                    for (int i = 0; i < T; ++i) {
                        const int c = j - i + T - 1;
                        acc[i][0] = fma_rn(ta[c], h.x, acc[i][0]);
                        acc[i][1] = fma_rn(ta[c], h.y, acc[i][1]);
                        acc[i][2] = fma_rn(tb[c], h.x, acc[i][2]);
                        acc[i][3] = fma_rn(tb[c], h.y, acc[i][3]);
                        acc[i][4] = fma_rn(tr[c], h.x, acc[i][4]);
                        acc[i][5] = fma_rn(tr[c], h.y, acc[i][5]);
                    }
                }
                ++hs;
            }
        }

        // ---- serial part: the recurrences, plus the <= T newest terms of each chain.
        // Shift-register form: before step i, acc[q] is the chain of output i+q; the new sample is
        // tap 64-q of that output.  After the step the finished chain acc[0] is consumed and the
        // register file shifts down by one (positions past T-1-i hold don't-care values).
#pragma unroll 1
        for (int i = 0; i < valid; ++i) {
            const float2 xv = fll_derotate(fs, agc_step(lc, cur[0], g));
            const float xr = xv.x, xi = xv.y;
            xs[(slot * T + i) * 32 + lane] = make_float2(xr, xi);
#pragma unroll
            for (int q = 0; q < T; ++q) {
                acc[q][0] = fma_rn(p.be_a[kHist - q], xr, acc[q][0]);
                acc[q][1] = fma_rn(p.be_a[kHist - q], xi, acc[q][1]);
                acc[q][2] = fma_rn(p.be_b[kHist - q], xr, acc[q][2]);
                acc[q][3] = fma_rn(p.be_b[kHist - q], xi, acc[q][3]);
                acc[q][4] = fma_rn(p.rrc[kHist - q], xr, acc[q][4]);
                acc[q][5] = fma_rn(p.rrc[kHist - q], xi, acc[q][5]);
            }
            fll_update<RE_ONLY>(lc, make_float2(acc[0][0], acc[0][1]), make_float2(acc[0][2], acc[0][3]), fs);
            // matched-filter output -> interpolator ring
            rs[((kITaps - 1 + n0 + i) & (RE - 1)) * 32 + lane] = make_float2(acc[0][4], acc[0][5]);
#pragma unroll
            for (int q = 0; q < T - 1; ++q) {
#pragma unroll
                for (int c = 0; c < 6; ++c) { acc[q][c] = acc[q + 1][c]; }
                cur[q] = cur[q + 1];
            }
        }

        // ---- symbols that became computable in this block (complex_fd.cpp:96 `while (offset < count)`)
        while (offset < n0 + valid) {
            const float2 y = timing_step<RE, 1>(kc, bank4, rs, lane, mu, om, offset);
            const float2 u = costas_loop_step(kc, y, cph, cfr, ph2);
            slicer_symbols<1>(p, ch, 1, sl, err_blocks, active, [&](int, bool) { return u; });
        }

        slot = (slot + 1 == S) ? 0 : slot + 1;
#pragma unroll
        for (int i = 0; i < T; ++i) { cur[i] = nxt[i]; }
    }

    // ---- carry state out
    if (active) {
        sp->agc_gain = g;
        fll_store(sp, fs);
        sp->tr_mu = mu; sp->tr_omega = om; sp->tr_offset = offset - count;   // complex_fd.cpp:145
        sp->costas_phase = cph; sp->costas_freq = cfr; sp->costas_ph2 = ph2;
        sp->n_samples += (unsigned long long)count;
        float2* xh = reinterpret_cast<float2*>(sp->x_hist);
        for (int m = 0; m < kHist; ++m) {
            const long long q = (long long)count + m;
            xh[m] = xs[(int)(((1 + q / T) % S) * T + (q % T)) * 32 + lane];
        }
        float2* rh = reinterpret_cast<float2*>(sp->r_hist);
        for (int j = 0; j < kITaps - 1; ++j) { rh[j] = rs[((count + j) & (RE - 1)) * 32 + lane]; }
    }
    slicer_store(p, sp, ch, sl, err_blocks, active);
}

template <int T>
int launch_tpc_t(const DemodParams& p_in, cudaStream_t stream, int warps_per_cta) {
    using L = TpcLayout<T>;
    DemodParams p = p_in;
    // tap tables padded with T-1 leading zeros (see the old-part loop)
    const float* src[3] = { p.be_a, p.be_b, p.rrc };
    for (int f = 0; f < 3; ++f) {
        for (int j = 0; j < kTapPad; ++j) {
            const int k = j - (T - 1);
            p.tpad[f][j] = (k >= 0 && k < kTaps) ? src[f][k] : 0.f;
        }
    }
    const int threads = 32 * warps_per_cta;
    const size_t smem = sizeof(float) * kIPhases * kITaps + L::kWarpBytes * warps_per_cta;
    const int grid = (p.n_channels + threads - 1) / threads;
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        kern<<<grid, threads, smem, stream>>>(p);
    };
    if (p.fastamp_re_only) { go(demod_tpc_kernel<T, true>); } else { go(demod_tpc_kernel<T, false>); }
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// ---------------------------------------------------------------------------------------
// 4-per-byte dibits <-> one dibit / one bit per byte, for the multi-GPU gather and the NETSYMS sink.
// ---------------------------------------------------------------------------------------
__global__ void pack_dibits_kernel(const uint8_t* __restrict__ dibits, long long in_stride,
                                   const int* __restrict__ counts, uint8_t* __restrict__ packed,
                                   long long out_stride, long long max_bytes, int n_channels) {
    for (int ch = blockIdx.y; ch < n_channels; ch += gridDim.y) {
        const int n = counts[ch];
        const uint8_t* src = dibits + (long long)ch * in_stride;
        uint8_t* dst = packed + (long long)ch * out_stride;
        for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < max_bytes;
             j += (long long)gridDim.x * blockDim.x) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long s = 4 * j + k;
                const uint32_t d = (s < n) ? (src[s] & 3u) : 0u;
                v |= d << (6 - 2 * k);
            }
            dst[j] = (uint8_t)v;
        }
    }
}

// packed rows -> the stream BitUnpacker::process writes (src/dsp/bit_unpacker.cpp:4-10: per dibit the high bit,
// then the low bit, one bit per byte) -- what the plugin's network sink sends (src/main.cpp:385-389) -- and/or
// one dibit per byte (DQPSKSymbolExtractor's output).  A thread expands 4 packed bytes (16 symbols): one 32-bit
// load, two 128-bit stores of bits / one of dibits.
__global__ void __launch_bounds__(256) unpack_dibits_kernel(const uint8_t* __restrict__ packed, long long in_stride,
                                                             const int* __restrict__ counts, uint8_t* __restrict__ dibits,
                                                             long long dibit_stride, uint8_t* __restrict__ bits, long long bit_stride,
                                                             int n_channels, long long max_words) {
    for (int ch = blockIdx.y; ch < n_channels; ch += gridDim.y) {
        const int n = counts[ch];
        const uint8_t* src = packed + (long long)ch * in_stride;
        const bool fast = ((in_stride & 3) == 0) && ((dibit_stride & 15) == 0) && ((bit_stride & 15) == 0) &&
                          ((reinterpret_cast<uintptr_t>(packed) & 3) == 0) && ((reinterpret_cast<uintptr_t>(dibits) & 15) == 0) &&
                          ((reinterpret_cast<uintptr_t>(bits) & 15) == 0);
        for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < max_words; w += (long long)gridDim.x * blockDim.x) {
            const long long s0 = 16 * w;                       // first symbol of this word
            if (s0 >= n) { break; }
            if (fast && s0 + 16 <= n) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(src + 4 * w);     // little endian: byte k = bits 8k..8k+7
                uint32_t d[4], bh[4], bl[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t b = (v >> (8 * k)) & 0xffu;
                    // dibits of byte b, first symbol in the lowest output byte
                    d[k] = ((b >> 6) & 3u) | (((b >> 4) & 3u) << 8) | (((b >> 2) & 3u) << 16) | ((b & 3u) << 24);
                    // bits: symbols 0,1 -> 4 bytes (hi0, lo0, hi1, lo1); symbols 2,3 -> the next 4
                    bh[k] = ((b >> 7) & 1u) | (((b >> 6) & 1u) << 8) | (((b >> 5) & 1u) << 16) | (((b >> 4) & 1u) << 24);
                    bl[k] = ((b >> 3) & 1u) | (((b >> 2) & 1u) << 8) | (((b >> 1) & 1u) << 16) | ((b & 1u) << 24);
                }
                if (dibits) { *reinterpret_cast<uint4*>(dibits + (long long)ch * dibit_stride + s0) = make_uint4(d[0], d[1], d[2], d[3]); }
                if (bits) {
                    uint4* o = reinterpret_cast<uint4*>(bits + (long long)ch * bit_stride + 2 * s0);
                    o[0] = make_uint4(bh[0], bl[0], bh[1], bl[1]);
                    o[1] = make_uint4(bh[2], bl[2], bh[3], bl[3]);
                }
            } else {
                for (long long s = s0; s < s0 + 16 && s < n; ++s) {
                    const uint32_t dv = (src[s >> 2] >> (6 - 2 * (int)(s & 3))) & 3u;
                    if (dibits) { dibits[(long long)ch * dibit_stride + s] = (uint8_t)dv; }
                    if (bits) {
                        bits[(long long)ch * bit_stride + 2 * s] = (uint8_t)(dv >> 1);
                        bits[(long long)ch * bit_stride + 2 * s + 1] = (uint8_t)(dv & 1u);
                    }
                }
            }
        }
    }
}

}  // namespace

int launch_tpc(const DemodParams& p, cudaStream_t stream, int T, int warps_per_cta) {
    return T == 4 ? launch_tpc_t<4>(p, stream, warps_per_cta) : launch_tpc_t<8>(p, stream, warps_per_cta);
}

const char* demod_variant_name(int variant) {
    switch (variant) {
        case 1: return "tpc4";
        case 2: return "tpc8";
        case 3: return "tpc4x4";
        case 4: return "ws4";
        case 5: return "ws4-p1";
        case 6: return "ws4-p2";
        case 7: return "ws4-p3";
        case 8: return "ws4-2cta";
        case 9: return "ws4-2cta-p1";
        case 10: return "ws4-2cta-p2";
        case 11: return "ws4-2cta-p3";
        default: return "auto";
    }
}

int demod_variant_count() { return kDemodVariants; }

int launch_demod(const DemodParams& p, int variant, cudaStream_t stream) {
    if (p.n_channels <= 0) { return 0; }
    // auto: the warp-specialised pipeline (ten role warps behind every 32 rows).  With one CTA per SM it has the
    // lowest latency per row; once there are more than 148 x 32 rows, two CTAs share an SM and fill each other's
    // stalls (the throughput regime: tdm_process_long_batch turns every recorded workload into this one).
    if (variant == 0) { variant = (p.n_channels > 148 * 32) ? 8 : 4; }
    switch (variant) {
        case 1: return launch_tpc(p, stream, 4, 1);
        case 2: return launch_tpc(p, stream, 8, 1);
        case 3: return launch_tpc(p, stream, 4, 4);
        case 4: case 5: case 6: case 7: return launch_ws4(p, stream, variant - 4, 1);
        case 8: case 9: case 10: case 11: return launch_ws4(p, stream, variant - 8, 2);
        default: return -1;
    }
}

int launch_pack_dibits(const uint8_t* dibits, long long in_stride, const int* counts, uint8_t* packed,
                       long long out_stride, int n_channels, long long max_syms, cudaStream_t stream) {
    if (n_channels <= 0) { return 0; }
    const long long max_bytes = (max_syms + 3) / 4;
    if (max_bytes <= 0) { return 0; }
    const int threads = 256;
    long long gx = (max_bytes + threads - 1) / threads;
    if (gx > 1024) { gx = 1024; }
    dim3 grid((unsigned)gx, (unsigned)(n_channels < 65535 ? n_channels : 65535));
    pack_dibits_kernel<<<grid, threads, 0, stream>>>(dibits, in_stride, counts, packed, out_stride, max_bytes, n_channels);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_unpack_dibits(const uint8_t* packed, long long in_stride, const int* counts, uint8_t* dibits, long long dibit_stride,
                         uint8_t* bits, long long bit_stride, int n_channels, long long max_syms, cudaStream_t stream) {
    if (n_channels <= 0 || max_syms <= 0) { return 0; }
    const long long max_words = (max_syms + 15) / 16;
    const int threads = 256;
    long long gx = (max_words + threads - 1) / threads;
    if (gx > 2048) { gx = 2048; }
    dim3 grid((unsigned)gx, (unsigned)(n_channels < 65535 ? n_channels : 65535));
    unpack_dibits_kernel<<<grid, threads, 0, stream>>>(packed, in_stride, counts, dibits, dibit_stride, bits, bit_stride, n_channels, max_words);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace tdm
