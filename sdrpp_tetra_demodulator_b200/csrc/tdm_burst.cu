// tdm_burst.cu -- burst synchroniser for C channels on sm_100a (C ABI: include/tdm_burst_b200.h).
//
// Reference behaviour (paths relative to the reference tree):
//   tetra_find_train_seq    src/decoder/src/phy/tetra_burst.c:271-341
//   tetra_burst_sync_in     src/decoder/src/phy/tetra_burst_sync.c:54-155   (+ make_bitbuf_space :38-51)
//   tetra_tdma_time_add_tn  src/decoder/src/tetra_tdma.c:44-74
//   _demodSinkHandler       src/main.cpp:385-414                            (training-sequence detector)
//
// The reference keeps a 4096-byte bit buffer per receiver, memmoves it on every call and memcmps byte
// strings.  A receiver's buffer is always a WINDOW of its bit stream (bits are appended at the end and dropped
// at the front), so here the stream is packed once, 1 bit per bit, and the buffer is two numbers: where the
// window starts and how long it is.  Three kernels per tdm_bsync_in:
//
//   pack    (HBM bound: reads 1 byte per input unit, writes 1/8 byte per bit)  -- every (channel, 4096-bit tile)
//           in parallel: 16-byte loads, 16 bytes -> 16 (or 32) bits with one multiply per 4 bytes, re-aligned
//           through shared memory behind the channel's carried bits;
//   match   (bit-parallel) -- every (channel, 32 positions): "does y / n-or-p start here" for 32 positions at once,
//           one funnel shift + one LOP3 per sequence bit, fully unrolled; writes two match bitmaps next to the
//           packed row.  With detect_ts it also matches the other five sequences and keeps the last hit per
//           channel for the src/main.cpp:385-414 detector;
//   sync    one warp per channel replays the reference's calls: the state machine is scalar (warp uniform),
//           a search is "first set bit of the match bitmap inside the buffer" (one word per lane + ballot), then
//           one 64-bit window compare to confirm type and remaining length.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>
#include "tdm_b200.h"
#include "tdm_burst_b200.h"
#include "tdm_internal.h"

namespace {

constexpr int kTileUnits = 16384;               // input bytes per pack CTA (16384 or 32768 bits): 4 16-byte loads in flight per thread
constexpr int kPadWords = 8;                    // readable words past the last bit of a row (64-bit window loads)

// training sequences as left-aligned numbers (first bit = most significant), phy/tetra_burst.c:61-72
constexpr unsigned long long seq_bits(const int* b, int n) {
    unsigned long long v = 0;
    for (int i = 0; i < n; ++i) { v = (v << 1) | (unsigned long long)b[i]; }
    return v;
}
constexpr int kN[22] = { 1,1, 0,1, 0,0, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,1, 0,0 };
constexpr int kP[22] = { 0,1, 1,1, 1,0, 1,0, 0,1, 0,0, 0,0, 1,1, 0,1, 1,1, 1,0 };
constexpr int kQ[22] = { 1,0, 1,1, 0,1, 1,1, 0,0, 0,0, 0,1, 1,0, 1,0, 1,1, 0,1 };
constexpr int kNN[33] = { 1,1,1, 0,0,1, 1,0,1, 1,1,1, 0,0,0, 1,1,1, 1,0,0, 0,1,1, 1,1,0, 0,0,0, 0,0,0 };
constexpr int kPP[33] = { 1,0,1, 0,1,1, 1,1,1, 1,0,1, 0,1,0, 1,0,1, 1,1,0, 0,0,1, 1,0,0, 0,1,0, 0,1,0 };
constexpr int kX[30] = { 1,0, 0,1, 1,1, 0,1, 0,0, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,1, 0,0, 0,0, 1,1 };
constexpr int kXX[45] = { 0,1,1,1,0,0,1,1,0,1,0,0,0,0,1,0,0,0,1,1,1,0,1,1,0,1,0,1,0,1,1,1,1,1,0,1,0,0,0,0,0,1,1,1,0 };
constexpr int kY[38] = { 1,1, 0,0, 0,0, 0,1, 1,0, 0,1, 1,1, 0,0, 1,1, 1,0, 1,0, 0,1, 1,1, 0,0, 0,0, 0,1, 1,0, 0,1, 1,1 };
constexpr unsigned long long kSeqN = seq_bits(kN, 22), kSeqP = seq_bits(kP, 22), kSeqQ = seq_bits(kQ, 22);
constexpr unsigned long long kSeqNN = seq_bits(kNN, 33), kSeqPP = seq_bits(kPP, 33), kSeqX = seq_bits(kX, 30);
constexpr unsigned long long kSeqXX = seq_bits(kXX, 45), kSeqY = seq_bits(kY, 38);
constexpr uint32_t kPreY = (uint32_t)(kSeqY >> 16), kPreX = (uint32_t)(kSeqX >> 8);   // first 22 bits

struct BsyncParams {
    const uint8_t* in;            // [C][in_stride] bytes
    long long in_stride;
    const int* n_units;           // [C] or null
    int units_all;
    int max_units;                // rows of wb are sized for this many units: per-channel counts are clamped to it
    int bits_per_unit;            // 1 or 2
    int n_channels;
    tdm_bsync_state* states;      // null for the stateless find
    uint32_t* wb;                 // [C][wstride] packed work rows: carried bits then the new bits
    uint32_t* my;                 // [C][wstride] bit (31 - q) of word w set <=> the SYNC sequence y starts at row position 32 w + q
    uint32_t* mnp;                // [C][wstride] same for the normal sequences n or p
    long long wstride;
    int* last_hit;                // [C] detector: 1 + index of the last new bit that completed a sequence, 0 = none
    int call_bits;
    tdm_burst* bursts;
    int max_bursts;
    int* n_bursts;
    int detect_ts;
    // stateless find
    uint32_t find_end, find_mask;
    int* find_type;
    uint32_t* find_offset;
};

__device__ __forceinline__ int units_of(const BsyncParams& p, int c) {
    int u = p.n_units ? p.n_units[c] : p.units_all;
    u = u < 0 ? 0 : u;
    return u > p.max_units ? p.max_units : u;
}

// ---------------------------------------------------------------------------------------------------
// pack: logical row = [carried bits (state.bitbuf, bits_in_buf of them)] ++ [new bits], MSB first.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack4_bits(uint32_t x) { return (((x & 0x01010101u) * 0x08040201u) >> 24) & 0xfu; }
__device__ __forceinline__ uint32_t pack4_dibits(uint32_t x) { return (((x & 0x03030303u) * 0x40100401u) >> 24) & 0xffu; }

// 16 input bytes around the ends of a row or from a row that is not 16-byte aligned: byte loads, zero outside
// [0, nu).  Deliberately NOT inlined: inlined, ptxas if-converts its 16 guarded loads into ~80 predicated
// instructions that every thread of the kernel issues (the first version of the pack kernel was issue bound at
// 0.37 warp instructions per input byte because of that).
__device__ __noinline__ uint4 load16_edge(const uint8_t* __restrict__ row, long long ua, int nu) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long u = ua + 4 * q + k;
            const uint32_t b = (u >= 0 && u < nu) ? (uint32_t)row[u] : 0u;
            acc |= b << (8 * k);
        }
        w[q] = acc;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) bsync_pack_kernel(const BsyncParams p) {
    constexpr int kGroups = kTileUnits / 16 + 1;                            // 16 units per group, one extra for the re-alignment
    __shared__ uint16_t s16[2 * kGroups + 6];
    const int c = blockIdx.y;
    const int bpu = p.bits_per_unit;
    const int tile_bits = kTileUnits * bpu, tile_words = tile_bits / 32;
    const int nu = units_of(p, c);
    const long long nbits = (long long)nu * bpu;
    const int cl = p.states ? (int)p.states[c].bits_in_buf : 0;
    const long long total = cl + nbits;
    const long long tile0 = (long long)blockIdx.x * tile_bits;              // first logical bit of this tile
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.last_hit) { p.last_hit[c] = 0; }
    if (tile0 >= total + 32 * kPadWords) { return; }
    const uint8_t* __restrict__ row = p.in + (long long)c * p.in_stride;
    const long long ib0 = tile0 - cl;                                       // first input bit of the tile (may be < 0)
    const long long u0 = ib0 >= 0 ? ib0 / bpu : -((-ib0 + bpu - 1) / bpu);  // floor(ib0 / bpu)
    const long long A = u0 & ~15LL;                                         // aligned first unit staged
    const bool aligned = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
    constexpr int kPer = (kGroups + 255) / 256;                             // groups per thread: all loads first, then the packing
    uint4 w[kPer];
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int g = threadIdx.x + 256 * r;
        const long long ua = A + 16LL * g;
        w[r] = make_uint4(0u, 0u, 0u, 0u);
        if (g < kGroups) {
            if (ua >= 0 && ua + 16 <= nu && aligned) { w[r] = __ldg(reinterpret_cast<const uint4*>(row + ua)); }
            else if (ua + 16 > 0 && ua < nu) { w[r] = load16_edge(row, ua, nu); }
        }
    }
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int g = threadIdx.x + 256 * r;
        if (g < kGroups) {
            if (bpu == 1) {
                s16[g] = (uint16_t)((pack4_bits(w[r].x) << 12) | (pack4_bits(w[r].y) << 8) | (pack4_bits(w[r].z) << 4) | pack4_bits(w[r].w));
            } else {
                reinterpret_cast<uint32_t*>(s16)[g] = ((pack4_dibits(w[r].z) << 24) | (pack4_dibits(w[r].w) << 16)) |
                                                      ((pack4_dibits(w[r].x) << 8) | pack4_dibits(w[r].y));     // s16[2g] = first 16 bits (little endian halves)
            }
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < tile_words; t += 256) {
        const long long wi = (long long)blockIdx.x * tile_words + t;
        if (wi * 32 < total + 32 * kPadWords && wi < p.wstride) {
            const int o = (int)(ib0 - A * bpu) + 32 * t;                    // bit offset into the staged bits
            const int idx = o >> 4, sh = o & 15;
            const unsigned long long v = ((unsigned long long)s16[idx] << 32) | ((unsigned long long)s16[idx + 1] << 16) | s16[idx + 2];
            uint32_t word = (uint32_t)(v >> (16 - sh));
            if (p.states && wi * 32 < cl) { word |= p.states[c].bitbuf[wi]; }   // carried bits (zero beyond bits_in_buf)
            p.wb[(long long)c * p.wstride + wi] = word;
        }
    }
}

// 64 bits of a packed row starting at bit position pos (first bit in bit 63)
__device__ __forceinline__ unsigned long long load64(const uint32_t* __restrict__ rowp, long long pos) {
    const long long wi = pos >> 5;
    const int sh = (int)(pos & 31);
    const uint32_t w0 = rowp[wi], w1 = rowp[wi + 1], w2 = rowp[wi + 2];
    const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
    return ((unsigned long long)hi << 32) | lo;
}

// A warp's window onto its packed row.  The sync kernel replays thousands of calls per channel and every call
// searches a few hundred bits right behind the previous call's: reading them from global memory made each of
// the ~8 search steps of a call wait for DRAM (5200 cycles per call measured).  The window is a shared-memory
// copy of kWinWords consecutive row words, refilled with coalesced loads only when a search leaves it
// (about every 11th call while locked).  With cache == nullptr it reads the row directly (stateless find).
constexpr int kWinWords = 512;                  // 16384 bits: the 4096-bit buffer + look-ahead, or one batch of 32 speculative calls
struct RowWindow {
    const uint32_t* __restrict__ rowp;          // packed bits
    const uint32_t* __restrict__ myp;           // match bitmaps (null for the stateless find)
    const uint32_t* __restrict__ mnpp;
    uint32_t* cache;                            // [3][kWinWords] shared memory of this warp (bits, my, mnp), or null
    long long cbase;                            // row word index of cache[0]; -1 = empty
    long long wstride;

    // make bits [pos, pos + nbits + 64) (and their bitmap words) readable
    __device__ __forceinline__ void ensure(long long pos, uint32_t nbits) {
        if (!cache) { return; }
        const long long w0 = pos >> 5, w1 = ((pos + nbits + 63) >> 5) + 2;
        if (cbase >= 0 && w0 >= cbase && w1 <= cbase + kWinWords) { return; }
        const int lane = threadIdx.x & 31;
        cbase = w0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < kWinWords / 32; ++k) {
            const long long wi = cbase + 32 * k + lane;
            const bool in = wi < wstride;
            cache[32 * k + lane] = in ? rowp[wi] : 0u;
            cache[kWinWords + 32 * k + lane] = in ? myp[wi] : 0u;
            cache[2 * kWinWords + 32 * k + lane] = in ? mnpp[wi] : 0u;
        }
        __syncwarp();
    }
    __device__ __forceinline__ unsigned long long get64(long long pos) const {
        if (!cache) { return load64(rowp, pos); }
        const int wi = (int)((pos >> 5) - cbase);
        const int sh = (int)(pos & 31);
        const uint32_t w0 = cache[wi], w1 = cache[wi + 1], w2 = cache[wi + 2];
        const uint32_t hi = __funnelshift_l(w1, w0, sh), lo = __funnelshift_l(w2, w1, sh);
        return ((unsigned long long)hi << 32) | lo;
    }
    // match-bitmap word wi of the row (absolute word index inside the window)
    __device__ __forceinline__ uint32_t mword(long long wi, bool with_np) const {
        const int k = (int)(wi - cbase);
        return cache[kWinWords + k] | (with_np ? cache[2 * kWinWords + k] : 0u);
    }
};

// which enabled sequence starts at a window W with `rem` bits left in the buffer; -1 = none.
// Order of the tests as in tetra_burst.c:309-338.
__device__ __forceinline__ int match_train_seq(unsigned long long W, uint32_t rem, uint32_t mask) {
    if ((mask & (1u << TDM_TRAIN_SYNC)) && rem >= 38 && (W >> 26) == kSeqY) { return TDM_TRAIN_SYNC; }
    if ((mask & (1u << TDM_TRAIN_NORM_1)) && rem >= 22 && (W >> 42) == kSeqN) { return TDM_TRAIN_NORM_1; }
    if ((mask & (1u << TDM_TRAIN_NORM_2)) && rem >= 22 && (W >> 42) == kSeqP) { return TDM_TRAIN_NORM_2; }
    if ((mask & (1u << TDM_TRAIN_NORM_3)) && rem >= 22 && (W >> 42) == kSeqQ) { return TDM_TRAIN_NORM_3; }
    if ((mask & (1u << TDM_TRAIN_EXT)) && rem >= 30 && (W >> 34) == kSeqX) { return TDM_TRAIN_EXT; }
    return -1;
}

// tetra_find_train_seq over buffer [ps, ps + len) of a packed row, by one warp.  `from` = first position >= ps + 21
// that still has to be examined with the plain rule (earlier ones are known not to hold an enabled sequence).
// Returns the type (warp uniform) and the offset relative to ps.
__device__ __forceinline__ int warp_find_train_seq(RowWindow& win, long long ps, uint32_t len, uint32_t mask,
                                                   long long from, uint32_t& offset) {
    const int lane = threadIdx.x & 31;
    if (len < 22) { return -1; }
    win.ensure(ps, len);
    // positions 0..20: the reference's look-ahead register does not hold in[i..i+21] there (one bit short preload,
    // tetra_burst.c:292-300): it holds in[i-1..19] ++ in[21..21+i] (a leading 0 for i = 0); the position is examined
    // only if THAT equals the first 22 bits of y, n, p, q or x.
    {
        const unsigned long long W0 = win.get64(ps);
        int t = -1;
        if (lane <= 20 && len - (uint32_t)lane >= 22) {
            const int i = lane;
            const uint32_t part1 = (uint32_t)(W0 >> 44) & ((1u << (21 - i)) - 1u);
            const uint32_t part2 = (uint32_t)(W0 >> (42 - i)) & ((1u << (i + 1)) - 1u);
            const uint32_t f = (part1 << (i + 1)) | part2;
            const bool pass = f == kPreY || f == (uint32_t)kSeqN || f == (uint32_t)kSeqP || f == (uint32_t)kSeqQ || f == kPreX;
            if (pass) { t = match_train_seq(win.get64(ps + i), len - (uint32_t)i, mask); }
        }
        const unsigned hit = __ballot_sync(0xffffffffu, t >= 0);
        if (hit) {
            const int src = __ffs(hit) - 1;
            offset = (uint32_t)src;
            return __shfl_sync(0xffffffffu, t, src);
        }
    }
    const long long pend = ps + (long long)len - 21;                        // positions with >= 22 bits left
    long long pbeg = ps + 21;
    if (from > pbeg) { pbeg = from; }
    for (long long pb = pbeg; pb < pend; pb += 32) {
        const long long pos = pb + lane;
        int t = -1;
        if (pos < pend) { t = match_train_seq(win.get64(pos), (uint32_t)(ps + len - pos), mask); }
        const unsigned hit = __ballot_sync(0xffffffffu, t >= 0);
        if (hit) {
            const int src = __ffs(hit) - 1;
            offset = (uint32_t)(pb + src - ps);
            return __shfl_sync(0xffffffffu, t, src);
        }
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------------
// match: for every row position, does a training sequence START there (content only; whether it still fits into
// the buffer is the state machine's business).  32 positions per thread: v_i = the 32 positions' i-th bits is one
// funnel shift of the aligned row words, and a sequence bit either keeps (AND) or clears (ANDN) candidates; after
// unrolling, the sequence is a compile-time constant, so each step is one SHF (shared by all sequences) + one LOP3.
//
// detect (src/main.cpp:385-414): new bit j completes a sequence iff D[j .. j+len) == seq with D = the previous 44
// bits followed by the new bits.  For j >= 44 the window starts at row position cl + j - 44, inside the new bits,
// so the same match words serve; the 44 positions that overlap the carried history are done one per thread.
// ---------------------------------------------------------------------------------------------------
template <int LEN>
__device__ __forceinline__ uint32_t match32(uint32_t a0, uint32_t a1, uint32_t a2, unsigned long long seq) {
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < LEN; ++i) {
        const uint32_t v = i == 0 ? a0 : (i < 32 ? __funnelshift_l(a1, a0, i) : (i == 32 ? a1 : __funnelshift_l(a2, a1, i - 32)));
        m &= ((seq >> (LEN - 1 - i)) & 1ull) ? v : ~v;
    }
    return m;
}

template <bool DETECT>
__global__ void __launch_bounds__(256) bsync_match_kernel(const BsyncParams p) {
    const int c = blockIdx.y;
    const long long n = (long long)units_of(p, c) * p.bits_per_unit;
    const int cl = (int)p.states[c].bits_in_buf;
    const long long total = cl + n;
    const uint32_t* __restrict__ rowp = p.wb + (long long)c * p.wstride;
    if (DETECT && blockIdx.x == 0 && threadIdx.x < 44 && (long long)threadIdx.x < n) {
        // window overlaps the carried 44-bit history: one position per thread, bit by bit
        const int j = threadIdx.x;
        const unsigned long long hist = ((unsigned long long)p.states[c].ts_window_hi << 32) | p.states[c].ts_window_lo;
        const unsigned long long fresh = load64(rowp, cl);
        unsigned long long w = 0;                                          // D[j .. j+45), first bit in bit 44; D[j+44] is new bit j
        for (int i = 0; i < 45; ++i) {
            const int d = j + i;
            const unsigned long long b = d < 44 ? (hist >> (43 - d)) & 1ull : (fresh >> (63 - (d - 44))) & 1ull;
            w = (w << 1) | b;
        }
        const bool hit = (w >> 23) == kSeqN || (w >> 23) == kSeqP || (w >> 23) == kSeqQ || (w >> 12) == kSeqNN || (w >> 12) == kSeqPP ||
                         (w >> 15) == kSeqX || w == kSeqXX || (w >> 7) == kSeqY;
        if (hit) { atomicMax(&p.last_hit[c], j + 1); }
    }
    const long long wi = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // positions 32 wi .. 32 wi + 31
    if (wi * 32 >= total || wi + 2 >= p.wstride) { return; }
    const uint32_t a0 = rowp[wi], a1 = rowp[wi + 1], a2 = rowp[wi + 2];
    const uint32_t my = match32<38>(a0, a1, a2, kSeqY);
    const uint32_t mnp = match32<22>(a0, a1, a2, kSeqN) | match32<22>(a0, a1, a2, kSeqP);
    p.my[(long long)c * p.wstride + wi] = my;
    p.mnp[(long long)c * p.wstride + wi] = mnp;
    if (DETECT) {
        uint32_t m = my | mnp | match32<22>(a0, a1, a2, kSeqQ) | match32<33>(a0, a1, a2, kSeqNN) | match32<33>(a0, a1, a2, kSeqPP) |
                     match32<30>(a0, a1, a2, kSeqX) | match32<45>(a0, a1, a2, kSeqXX);
        // window start L = 32 wi + q completes at new bit j = L - cl + 44: it counts iff L >= cl and j < n
        const long long lo = cl, hi = total - 44;                          // valid L in [lo, hi)
        const long long L0 = wi * 32;
        if (L0 + 32 <= lo || L0 >= hi) { m = 0; }
        else {
            if (L0 < lo) { m &= 0xffffffffu >> (int)(lo - L0); }
            if (L0 + 32 > hi) { m &= ~(0xffffffffu >> (int)(hi - L0)); }
        }
        if (m) {
            const int q = 31 - (__ffs(m) - 1);                              // the LAST matching position of the word
            atomicMax(&p.last_hit[c], (int)(L0 + q - cl + 44) + 1);
        }
    }
}

// The state machine's search: tetra_find_train_seq over buffer [ps, ps + len) with mask {SYNC} (with_np = false) or
// {NORM_1, NORM_2, SYNC}, driven by the match bitmaps.  `from` as in warp_find_train_seq.
__device__ __forceinline__ int warp_find_bitmap(RowWindow& win, long long ps, uint32_t len, bool with_np, long long from, uint32_t& offset) {
    const int lane = threadIdx.x & 31;
    if (len < 22) { return -1; }
    win.ensure(ps, len);
    const uint32_t mask = with_np ? ((1u << TDM_TRAIN_NORM_1) | (1u << TDM_TRAIN_NORM_2) | (1u << TDM_TRAIN_SYNC)) : (1u << TDM_TRAIN_SYNC);
    // positions 0..20 (look-ahead quirk, see warp_find_train_seq): only worth a look if a sequence starts there at all
    {
        const long long w0 = ps >> 5;
        const int sh = (int)(ps & 31);
        const uint32_t head = __funnelshift_l(win.mword(w0 + 1, with_np), win.mword(w0, with_np), sh) >> 11;    // positions ps .. ps+20
        if (head) {
            const unsigned long long W0 = win.get64(ps);
            int t = -1;
            if (lane <= 20 && len - (uint32_t)lane >= 22) {
                const int i = lane;
                const uint32_t part1 = (uint32_t)(W0 >> 44) & ((1u << (21 - i)) - 1u);
                const uint32_t part2 = (uint32_t)(W0 >> (42 - i)) & ((1u << (i + 1)) - 1u);
                const uint32_t f = (part1 << (i + 1)) | part2;
                const bool pass = f == kPreY || f == (uint32_t)kSeqN || f == (uint32_t)kSeqP || f == (uint32_t)kSeqQ || f == kPreX;
                if (pass) { t = match_train_seq(win.get64(ps + i), len - (uint32_t)i, mask); }
            }
            const unsigned hit = __ballot_sync(0xffffffffu, t >= 0);
            if (hit) {
                const int src = __ffs(hit) - 1;
                offset = (uint32_t)src;
                return __shfl_sync(0xffffffffu, t, src);
            }
        }
    }
    long long lo = ps + 21;
    if (from > lo) { lo = from; }
    const long long hi = ps + (long long)len - 21;                          // positions with >= 22 bits left
    for (long long wbase = lo >> 5; wbase * 32 < hi; wbase += 32) {
        const long long wi = wbase + lane;
        const long long L0 = wi * 32;
        uint32_t m = 0;
        if (L0 < hi && L0 + 32 > lo) {
            m = win.mword(wi, with_np);
            if (L0 < lo) { m &= 0xffffffffu >> (int)(lo - L0); }
            if (L0 + 32 > hi) { m &= ~(0xffffffffu >> (int)(hi - L0)); }
        }
        for (;;) {
            const unsigned cand = __ballot_sync(0xffffffffu, m != 0u);
            if (!cand) { break; }
            const int src = __ffs(cand) - 1;
            const int q = __clz(__shfl_sync(0xffffffffu, m, src));          // earliest position = most significant set bit
            const long long pos = (wbase + src) * 32 + q;
            const int t = match_train_seq(win.get64(pos), (uint32_t)(ps + len - pos), mask);
            if (t >= 0) { offset = (uint32_t)(pos - ps); return t; }
            if (lane == src) { m &= ~(0x80000000u >> q); }                  // y with fewer than 38 bits left: not a match yet
        }
    }
    return -1;
}

// warp_find_bitmap for ONE lane working on its own buffer (mask {NORM_1, NORM_2, SYNC}); the window must already
// cover [ps, ps + len + 64).  Used by the speculative slot-parallel path of the sync kernel.
__device__ __forceinline__ int lane_find_bitmap(const RowWindow& win, long long ps, uint32_t len, uint32_t& offset) {
    const uint32_t mask = (1u << TDM_TRAIN_NORM_1) | (1u << TDM_TRAIN_NORM_2) | (1u << TDM_TRAIN_SYNC);
    if (len < 22) { return -1; }
    {
        const long long w0 = ps >> 5;
        const int sh = (int)(ps & 31);
        uint32_t head = __funnelshift_l(win.mword(w0 + 1, true), win.mword(w0, true), sh) >> 11;    // positions ps .. ps+20, ps in bit 20
        if (head) {
            const unsigned long long W0 = win.get64(ps);
            while (head) {
                const int i = 20 - (31 - __clz(head));
                head &= ~(1u << (20 - i));
                if (len - (uint32_t)i < 22) { break; }
                const uint32_t part1 = (uint32_t)(W0 >> 44) & ((1u << (21 - i)) - 1u);
                const uint32_t part2 = (uint32_t)(W0 >> (42 - i)) & ((1u << (i + 1)) - 1u);
                const uint32_t f = (part1 << (i + 1)) | part2;
                if (f == kPreY || f == (uint32_t)kSeqN || f == (uint32_t)kSeqP || f == (uint32_t)kSeqQ || f == kPreX) {
                    const int t = match_train_seq(win.get64(ps + i), len - (uint32_t)i, mask);
                    if (t >= 0) { offset = (uint32_t)i; return t; }
                }
            }
        }
    }
    const long long lo = ps + 21, hi = ps + (long long)len - 21;
    for (long long wi = lo >> 5; wi * 32 < hi; ++wi) {
        const long long L0 = wi * 32;
        uint32_t m = win.mword(wi, true);
        if (L0 < lo) { m &= 0xffffffffu >> (int)(lo - L0); }
        if (L0 + 32 > hi) { m &= ~(0xffffffffu >> (int)(hi - L0)); }
        while (m) {
            const int q = __clz(m);
            const long long pos = L0 + q;
            const int t = match_train_seq(win.get64(pos), (uint32_t)(ps + len - pos), mask);
            if (t >= 0) { offset = (uint32_t)(pos - ps); return t; }
            m &= ~(0x80000000u >> q);
        }
    }
    return -1;
}

// m >= 1 calls of tetra_tdma_time_add_tn(&time, 1) in closed form (tetra_tdma.c:27-74: each counter wraps from its
// maximum back to 1 and carries one into the next; a counter that still holds its initial 0 simply counts up).
__device__ __forceinline__ void advance_time(uint32_t& tn, uint32_t& fn, uint32_t& mn, uint32_t m) {
    if (m == 0 || tn > 4 || fn > 18 || mn > 60) {                           // states the reference itself never produces: step by step
        for (uint32_t i = 0; i < m; ++i) {
            tn += 1;
            if (tn > 4) { const uint32_t d = tn / 4; tn %= 4; fn += d; }
            if (fn > 18) { const uint32_t d = fn / 18; fn %= 18; mn += d; }
            if (mn > 60) { mn %= 60; }
        }
        return;
    }
    const int t = (int)tn - 1 + (int)m;                                    // >= 0
    const uint32_t w1 = (uint32_t)(t / 4);
    tn = (uint32_t)(t % 4) + 1u;
    if (w1 == 0) { return; }
    const int f = (int)fn - 1 + (int)w1;
    const uint32_t w2 = (uint32_t)(f / 18);
    fn = (uint32_t)(f % 18) + 1u;
    if (w2 == 0) { return; }
    const int q = (int)mn - 1 + (int)w2;
    mn = (uint32_t)(q % 60) + 1u;
}

// ---------------------------------------------------------------------------------------------------
// sync: one warp per channel replays tetra_burst_sync_in call by call.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bsync_fsm_kernel(const BsyncParams p) {
    __shared__ uint32_t win_s[4][3 * kWinWords];
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= p.n_channels) { return; }
    tdm_bsync_state* __restrict__ sp = p.states + c;
    const uint32_t* __restrict__ rowp = p.wb + (long long)c * p.wstride;
    const uint32_t n = (uint32_t)units_of(p, c) * (uint32_t)p.bits_per_unit;
    int state = sp->state;
    uint32_t bib = sp->bits_in_buf;
    const uint32_t cl0 = bib;
    const uint32_t base = sp->bitbuf_start_bitnum;                          // row position 0 <-> this stream position
    uint32_t start = base;
    uint32_t nfs = sp->next_frame_start_bitnum;
    uint32_t tn = sp->tn, fn = sp->fn, mn = sp->mn;
    uint32_t cursor = sp->searched_upto;
    if ((int32_t)(cursor - base) < 0) { cursor = base; }
    RowWindow win{ rowp, p.my + (long long)c * p.wstride, p.mnp + (long long)c * p.wstride, win_s[threadIdx.x >> 5], -1, p.wstride };

    if (p.detect_ts && lane == 0) {
        // finish the detector: expiry counter and the newest 44 bits (src/main.cpp:403-412)
        uint32_t found = sp->ts_found, expire = sp->ts_expire;
        const int last = p.last_hit[c];
        if (last > 0) {
            const long long e = 2048LL - ((long long)n - (last - 1));
            if (e > 0) { found = 1; expire = (uint32_t)e; } else { found = 0; expire = 0; }
        } else if (expire > 0) {
            if (expire <= n) { found = 0; expire = 0; } else { expire -= n; }
        }
        unsigned long long hist = ((unsigned long long)sp->ts_window_hi << 32) | sp->ts_window_lo;
        if (n >= 44) { hist = load64(rowp, (long long)cl0 + n - 44) >> 20; }
        else if (n > 0) { hist = ((hist << n) | (load64(rowp, cl0) >> (64 - n))) & ((1ull << 44) - 1ull); }
        sp->ts_found = found; sp->ts_expire = expire;
        sp->ts_window_lo = (uint32_t)hist; sp->ts_window_hi = (uint32_t)(hist >> 32);
    }

    uint32_t nb = 0, call = 0;
    const uint32_t call_bits = (uint32_t)p.call_bits;
    uint32_t off = 0;
    while (off < n) {
        if (state == TDM_RX_S_LOCKED && bib + call_bits <= TDM_BSYNC_BITBUF) {
            // ---- LOCKED: slot-parallel replay.  While the receiver stays LOCKED, what call k sees is known in
            // advance: the buffer gains call_bits per call and loses one slot per call whenever it holds 510 bits, so
            // with a_k = bib + (bits of calls 0..k), S_k = min(k + 1, floor(a_k / 510)) slots are gone after call k
            // (a queue served at most once per step whose arrivals complete at most one slot per step; the buffer
            // cannot overflow because it never grows while it holds a slot).  Lane k replays call k on its own; the
            // first lane that would leave LOCKED ends the batch, later lanes are discarded and replayed.
            const uint32_t remaining = n - off;
            const uint32_t ncalls = (remaining + call_bits - 1) / call_bits;
            // the buffers of all calls of a batch must fit into the window together
            const uint32_t fit = ((uint32_t)(32 * kWinWords - 200) - bib) / call_bits;
            const uint32_t B = min(min(32u, ncalls), fit);                  // >= 1: bib + call_bits <= 4096
            const long long ps0 = (long long)(start - base);
            const uint32_t span = B * call_bits < remaining ? B * call_bits : remaining;
            win.ensure(ps0, bib + span);
            const bool active = (uint32_t)lane < B;
            const uint32_t arr_prev = min((uint32_t)lane * call_bits, remaining);
            const uint32_t arr = min(((uint32_t)lane + 1u) * call_bits, remaining);
            const uint32_t cons_before = min((uint32_t)lane, (bib + arr_prev) / TDM_BITS_PER_TS);
            const uint32_t cons_after = min((uint32_t)lane + 1u, (bib + arr) / TDM_BITS_PER_TS);
            const bool proc = active && cons_after > cons_before;          // this call finds a whole slot in the buffer
            const uint32_t bibf = bib + arr - TDM_BITS_PER_TS * cons_before;
            const long long psk = ps0 + (long long)TDM_BITS_PER_TS * cons_before;
            uint32_t offs = 0;
            int rc = -1;
            if (proc) { rc = lane_find_bitmap(win, psk, bibf, offs); }
            const bool is_sync = rc == TDM_TRAIN_SYNC, is_norm = rc == TDM_TRAIN_NORM_1 || rc == TDM_TRAIN_NORM_2;
            const bool deliver = proc && ((is_sync && offs == 214) || (is_norm && offs == 244));
            const bool unlock = proc && !deliver && !is_norm;              // SYNC at the wrong place, or nothing found
            const unsigned ub = __ballot_sync(0xffffffffu, unlock);
            const uint32_t last = ub ? (uint32_t)(__ffs(ub) - 1) : B - 1u;  // last call of the batch that really happens
            const bool committed = active && (uint32_t)lane <= last;
            const unsigned db = __ballot_sync(0xffffffffu, committed && deliver);
            if (committed && deliver) {
                const uint32_t idx = nb + (uint32_t)__popc(db & ((1u << lane) - 1u));
                if (idx < (uint32_t)p.max_bursts && p.bursts) {
                    tdm_burst* b = p.bursts + ((long long)c * p.max_bursts + idx);
                    uint32_t t1 = tn, f1 = fn, m1 = mn;
                    advance_time(t1, f1, m1, cons_after);
                    b->bitnum = start + TDM_BITS_PER_TS * cons_before; b->train_seq = rc; b->tn = t1; b->fn = f1; b->mn = m1;
                    b->call_index = call + (uint32_t)lane; b->reserved[0] = 0; b->reserved[1] = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        uint32_t v = (uint32_t)(win.get64(psk + 32 * k) >> 32);
                        if (k == 15) { v &= 0xfffffffcu; }
                        b->bits[k] = v;
                    }
                }
            }
            const uint32_t arr_c = min((last + 1u) * call_bits, remaining);
            const uint32_t cons_c = min(last + 1u, (bib + arr_c) / TDM_BITS_PER_TS);
            advance_time(tn, fn, mn, cons_c);
            bib = bib + arr_c - TDM_BITS_PER_TS * cons_c;
            start += TDM_BITS_PER_TS * cons_c; nfs += TDM_BITS_PER_TS * cons_c;
            nb += (uint32_t)__popc(db);
            off += arr_c; call += last + 1u;
            if (ub) { state = TDM_RX_S_UNLOCKED; cursor = start; }
            continue;
        }
        // ---- one call, the whole warp on it (acquisition, and LOCKED with more than a slot buffered)
        const uint32_t len = n - off < call_bits ? n - off : call_bits;
        off += len;
        const uint32_t this_call = call++;
        do {
            const uint32_t space = TDM_BSYNC_BITBUF - bib;
            if (space < len) { const uint32_t delta = len - space; bib -= delta; start += delta; }   // make_bitbuf_space
            bib += len;
            const long long ps = (long long)(start - base);
            uint32_t offs = 0;
            if (state == TDM_RX_S_UNLOCKED) {
                if (bib < 2 * TDM_BITS_PER_TS) { break; }
                if ((int32_t)(cursor - start) < 0) { cursor = start; }
                const int rc = warp_find_bitmap(win, ps, bib, false, (long long)(cursor - base), offs);
                if (rc < 0) {
                    cursor = start + bib - 37;       // positions up to end - 38 had all 38 bits and did not hold y
                    break;
                }
                state = TDM_RX_S_KNOW_FSTART;
                nfs = start + offs + 296;
                break;
            }
            if (state == TDM_RX_S_KNOW_FSTART) {
                if (start + bib < nfs) { break; }
                uint32_t shift = nfs - start;
                if ((int32_t)shift < 0) { shift = 0; }   // undefined in the reference; see tdm_burst_b200.h
                bib -= shift; start += shift;
                nfs += TDM_BITS_PER_TS;
                state = TDM_RX_S_LOCKED;             // falls through into the LOCKED case, like the reference
            }
            if (bib < TDM_BITS_PER_TS) { break; }
            advance_time(tn, fn, mn, 1);             // tetra_tdma_time_add_tn(&time, 1)
            const long long ps2 = (long long)(start - base);
            const int rc = warp_find_bitmap(win, ps2, bib, true, ps2, offs);
            bool deliver = false;
            if (rc == TDM_TRAIN_SYNC) {
                if (offs == 214) { deliver = true; } else { state = TDM_RX_S_UNLOCKED; cursor = start + TDM_BITS_PER_TS; }
            } else if (rc == TDM_TRAIN_NORM_1 || rc == TDM_TRAIN_NORM_2) {
                deliver = (offs == 244);
            } else {
                state = TDM_RX_S_UNLOCKED; cursor = start + TDM_BITS_PER_TS;
            }
            if (deliver) {
                if (nb < (uint32_t)p.max_bursts && p.bursts) {
                    tdm_burst* b = p.bursts + ((long long)c * p.max_bursts + nb);
                    if (lane == 16) {
                        b->bitnum = start; b->train_seq = rc; b->tn = tn; b->fn = fn; b->mn = mn; b->call_index = this_call;
                        b->reserved[0] = 0; b->reserved[1] = 0;
                    }
                    if (lane < 16) {                                                // the burst is already packed: 16 words
                        uint32_t v = (uint32_t)(win.get64(ps2 + 32 * lane) >> 32);
                        if (lane == 15) { v &= 0xfffffffcu; }                       // bits 510, 511 do not exist
                        b->bits[lane] = v;
                    }
                }
                ++nb;
            }
            bib -= TDM_BITS_PER_TS; start += TDM_BITS_PER_TS; nfs += TDM_BITS_PER_TS;
        } while (0);
    }

    // carry the buffered bits [start, start + bib) to the front of the state's packed buffer
    {
        const long long ps = (long long)(start - base);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t w = 4 * lane + k;
            uint32_t v = 0;
            if (32 * w < bib) {
                v = (uint32_t)(load64(rowp, ps + 32 * w) >> 32);
                const uint32_t left = bib - 32 * w;
                if (left < 32) { v &= ~(0xffffffffu >> left); }
            }
            sp->bitbuf[w] = v;
        }
    }
    if (lane == 0) {
        sp->state = state; sp->bits_in_buf = bib; sp->bitbuf_start_bitnum = start; sp->next_frame_start_bitnum = nfs;
        sp->tn = tn; sp->fn = fn; sp->mn = mn; sp->searched_upto = cursor;
        sp->n_bits += n; sp->n_bursts += nb;
        if (p.n_bursts) { p.n_bursts[c] = (int)nb; }
    }
}

__global__ void __launch_bounds__(128) bsync_find_kernel(const BsyncParams p) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= p.n_channels) { return; }
    uint32_t offs = 0;
    RowWindow win{ p.wb + (long long)c * p.wstride, nullptr, nullptr, nullptr, -1, p.wstride };
    const int rc = warp_find_train_seq(win, 0, p.find_end, p.find_mask, 0, offs);
    if (lane == 0) { p.find_type[c] = rc; p.find_offset[c] = rc >= 0 ? offs : 0u; }
}

long long words_for(long long bits) { return (bits + 31) / 32 + kPadWords; }

}  // namespace

struct tdm_bsync {
    int device = 0;
    int n_channels = 0;
    long long max_units = 0;
    long long wstride = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    long long launches = 0;
    tdm_bsync_state* d_states = nullptr;
    uint32_t* d_wb = nullptr;
    int* d_last_hit = nullptr;
    // staging for TDM_MEM_HOST callers
    uint8_t* d_in = nullptr;
    int* d_units = nullptr;
    int* d_nbursts = nullptr;
    tdm_burst* d_bursts = nullptr;
    int d_bursts_cap = 0;
    cudaEvent_t ev[4] = {};             // around pack / detect / sync of the most recent tdm_bsync_in
    bool ev_detect = false;
};

#define BS_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) { return tdm_internal_fail(TDM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); } \
    } while (0)

namespace {
struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) { cudaSetDevice(dev); } else { prev = -1; } }
    ~DevGuard() { if (prev >= 0) { cudaSetDevice(prev); } }
};

int check_device(const char* who, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        return tdm_internal_fail(TDM_ERR_NO_DEVICE, "%s: no CUDA device (this library has no CPU fallback)", who);
    }
    if (device < 0 || device >= ndev) { return tdm_internal_fail(TDM_ERR_ARG, "%s: device %d out of range (%d devices)", who, device, ndev); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        return tdm_internal_fail(TDM_ERR_NO_DEVICE, "%s: device %d is not sm_100; this library is built for sm_100a only", who, device);
    }
    return TDM_OK;
}
}  // namespace

extern "C" {

int tdm_bsync_create(int32_t n_channels, int64_t max_units, int32_t device, tdm_bsync** out) {
    if (!out) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_create: out is null"); }
    *out = nullptr;
    if (n_channels <= 0 || max_units <= 0 || max_units > (1LL << 29)) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_create: n_channels and max_units must be > 0 (max_units <= 2^29)"); }
    if (n_channels > 65535) { return tdm_internal_fail(TDM_ERR_UNSUPPORTED, "tdm_bsync_create: at most 65535 channels per handle (one grid row per channel); use several handles"); }
    int rc = check_device("tdm_bsync_create", device);
    if (rc != TDM_OK) { return rc; }
    tdm_bsync* h = new (std::nothrow) tdm_bsync();
    if (!h) { return tdm_internal_fail(TDM_ERR_NOMEM, "tdm_bsync_create: out of host memory"); }
    h->device = device; h->n_channels = n_channels; h->max_units = max_units;
    h->wstride = words_for(TDM_BSYNC_BITBUF + 2 * max_units);
    DevGuard guard(device);
    auto cleanup = [&](int code) { tdm_bsync_destroy(h); return code; };
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { return cleanup(tdm_internal_fail(TDM_ERR_CUDA, "cudaStreamCreate failed")); }
    h->stream = h->own_stream;
    for (auto& e : h->ev) { if (cudaEventCreate(&e) != cudaSuccess) { return cleanup(tdm_internal_fail(TDM_ERR_CUDA, "cudaEventCreate failed")); } }
    if (cudaMalloc(&h->d_states, sizeof(tdm_bsync_state) * (size_t)n_channels) != cudaSuccess ||
        cudaMalloc(&h->d_wb, 3 * sizeof(uint32_t) * (size_t)n_channels * (size_t)h->wstride) != cudaSuccess ||
        cudaMalloc(&h->d_last_hit, sizeof(int) * (size_t)n_channels) != cudaSuccess) {
        return cleanup(tdm_internal_fail(TDM_ERR_NOMEM, "tdm_bsync_create: cudaMalloc failed"));
    }
    rc = tdm_bsync_reset(h);
    if (rc != TDM_OK) { return cleanup(rc); }
    *out = h;
    return TDM_OK;
}

int tdm_bsync_destroy(tdm_bsync* h) {
    if (!h) { return TDM_OK; }
    DevGuard guard(h->device);
    if (h->own_stream) { cudaStreamSynchronize(h->own_stream); }
    cudaFree(h->d_states); cudaFree(h->d_wb); cudaFree(h->d_last_hit); cudaFree(h->d_in); cudaFree(h->d_units);
    cudaFree(h->d_nbursts); cudaFree(h->d_bursts);
    for (auto& e : h->ev) { if (e) { cudaEventDestroy(e); } }
    if (h->own_stream) { cudaStreamDestroy(h->own_stream); }
    delete h;
    return TDM_OK;
}

int tdm_bsync_set_stream(tdm_bsync* h, void* cuda_stream) {
    if (!h) { return tdm_internal_fail(TDM_ERR_ARG, "null handle"); }
    h->stream = (cuda_stream == TDM_OWN_STREAM) ? h->own_stream : (cudaStream_t)cuda_stream;
    return TDM_OK;
}

int tdm_bsync_reset(tdm_bsync* h) {
    if (!h) { return tdm_internal_fail(TDM_ERR_ARG, "null handle"); }
    DevGuard guard(h->device);
    BS_CUDA(cudaMemsetAsync(h->d_states, 0, sizeof(tdm_bsync_state) * (size_t)h->n_channels, h->stream));   // talloc_zero
    BS_CUDA(cudaStreamSynchronize(h->stream));
    return TDM_OK;
}

int tdm_bsync_get_state(tdm_bsync* h, tdm_bsync_state* host_states, int32_t n_channels) {
    if (!h || !host_states || n_channels != h->n_channels) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_get_state: bad arguments"); }
    DevGuard guard(h->device);
    BS_CUDA(cudaMemcpyAsync(host_states, h->d_states, sizeof(tdm_bsync_state) * (size_t)n_channels, cudaMemcpyDeviceToHost, h->stream));
    BS_CUDA(cudaStreamSynchronize(h->stream));
    return TDM_OK;
}

int tdm_bsync_set_state(tdm_bsync* h, const tdm_bsync_state* host_states, int32_t n_channels) {
    if (!h || !host_states || n_channels != h->n_channels) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_set_state: bad arguments"); }
    for (int c = 0; c < n_channels; ++c) {
        if (host_states[c].bits_in_buf > TDM_BSYNC_BITBUF || host_states[c].state < 0 || host_states[c].state > TDM_RX_S_LOCKED) {
            return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_set_state: channel %d holds an impossible state", c);
        }
    }
    DevGuard guard(h->device);
    BS_CUDA(cudaMemcpyAsync(h->d_states, host_states, sizeof(tdm_bsync_state) * (size_t)n_channels, cudaMemcpyHostToDevice, h->stream));
    BS_CUDA(cudaStreamSynchronize(h->stream));
    return TDM_OK;
}

int64_t tdm_bsync_launch_count(const tdm_bsync* h) { return h ? h->launches : 0; }

int tdm_bsync_last_kernel_ms(tdm_bsync* h, float* ms3) {
    if (!h || !ms3) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_last_kernel_ms: null argument"); }
    if (h->launches == 0) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_last_kernel_ms: nothing launched yet"); }
    DevGuard guard(h->device);
    BS_CUDA(cudaEventSynchronize(h->ev[3]));
    BS_CUDA(cudaEventElapsedTime(&ms3[0], h->ev[0], h->ev[1]));
    ms3[1] = 0.f;
    if (h->ev_detect) { BS_CUDA(cudaEventElapsedTime(&ms3[1], h->ev[1], h->ev[2])); }
    BS_CUDA(cudaEventElapsedTime(&ms3[2], h->ev[2], h->ev[3]));
    return TDM_OK;
}

int tdm_bsync_in(tdm_bsync* h, const uint8_t* in, int64_t in_stride, const int32_t* n_units, int32_t units_all,
                 int32_t in_kind, int32_t call_bits, tdm_burst* bursts, int32_t max_bursts, int32_t* n_bursts,
                 int32_t detect_ts, int32_t mem_kind) {
    if (!h) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: null handle"); }
    if (in_kind != TDM_BSYNC_IN_BITS && in_kind != TDM_BSYNC_IN_DIBITS) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: in_kind"); }
    if (call_bits < 1 || call_bits > TDM_BSYNC_MAX_CALL_BITS) {
        return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: call_bits must be 1..%d (the reference is undefined beyond one slot per call)", TDM_BSYNC_MAX_CALL_BITS);
    }
    if (max_bursts < 0 || (max_bursts > 0 && !bursts)) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: bursts"); }
    if (mem_kind != TDM_MEM_HOST && mem_kind != TDM_MEM_DEVICE) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: mem_kind"); }
    const int C = h->n_channels;
    long long umax = units_all;
    if (n_units && mem_kind == TDM_MEM_HOST) {
        umax = 0;
        for (int c = 0; c < C; ++c) { if (n_units[c] > umax) { umax = n_units[c]; } }
    } else if (n_units) {
        umax = h->max_units;                        // per-channel counts live on the device: the grid covers the handle's maximum
    }
    if (umax < 0 || umax > h->max_units) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: %lld units per channel > max_units %lld", umax, h->max_units); }
    // (per-channel counts that live on the device cannot be checked here: the caller guarantees n_units[c] <= in_stride)
    const bool counts_on_device = n_units && mem_kind == TDM_MEM_DEVICE;
    if (umax > 0 && (!in || (!counts_on_device && in_stride < umax))) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: input pointer / stride"); }
    if (mem_kind == TDM_MEM_DEVICE && bursts && (reinterpret_cast<uintptr_t>(bursts) & 3u)) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_bsync_in: bursts must be 4-byte aligned"); }
    DevGuard guard(h->device);

    BsyncParams p{};
    p.in_stride = in_stride; p.units_all = units_all; p.bits_per_unit = in_kind == TDM_BSYNC_IN_DIBITS ? 2 : 1;
    p.max_units = (int)h->max_units;
    p.n_channels = C; p.states = h->d_states; p.wb = h->d_wb; p.wstride = h->wstride; p.last_hit = h->d_last_hit;
    p.my = h->d_wb + (size_t)C * (size_t)h->wstride; p.mnp = p.my + (size_t)C * (size_t)h->wstride;
    p.call_bits = call_bits; p.max_bursts = max_bursts; p.detect_ts = detect_ts ? 1 : 0;
    if (mem_kind == TDM_MEM_DEVICE) {
        p.in = in; p.n_units = n_units; p.bursts = bursts; p.n_bursts = n_bursts;
    } else {
        if (!h->d_in) { BS_CUDA(cudaMalloc(&h->d_in, (size_t)C * (size_t)h->max_units)); }
        if (!h->d_units) { BS_CUDA(cudaMalloc(&h->d_units, sizeof(int) * (size_t)C)); }
        if (!h->d_nbursts) { BS_CUDA(cudaMalloc(&h->d_nbursts, sizeof(int) * (size_t)C)); }
        if (max_bursts > h->d_bursts_cap) {
            cudaFree(h->d_bursts); h->d_bursts = nullptr; h->d_bursts_cap = 0;
            BS_CUDA(cudaMalloc(&h->d_bursts, sizeof(tdm_burst) * (size_t)C * (size_t)max_bursts));
            h->d_bursts_cap = max_bursts;
        }
        if (umax > 0) {
            BS_CUDA(cudaMemcpy2DAsync(h->d_in, (size_t)h->max_units, in, (size_t)in_stride, (size_t)umax, (size_t)C, cudaMemcpyHostToDevice, h->stream));
        }
        if (n_units) { BS_CUDA(cudaMemcpyAsync(h->d_units, n_units, sizeof(int) * (size_t)C, cudaMemcpyHostToDevice, h->stream)); }
        p.in = h->d_in; p.in_stride = h->max_units; p.n_units = n_units ? h->d_units : nullptr;
        p.bursts = h->d_bursts; p.n_bursts = h->d_nbursts;
    }
    const long long row_bits = TDM_BSYNC_BITBUF + umax * p.bits_per_unit + 32 * kPadWords;
    const long long tile_bits = (long long)kTileUnits * p.bits_per_unit;
    dim3 gpack((unsigned)((row_bits + tile_bits - 1) / tile_bits), (unsigned)C);
    cudaEventRecord(h->ev[0], h->stream);
    bsync_pack_kernel<<<gpack, 256, 0, h->stream>>>(p);
    h->launches++;
    cudaEventRecord(h->ev[1], h->stream);
    h->ev_detect = true;
    {
        const long long words = (TDM_BSYNC_BITBUF + umax * p.bits_per_unit + 31) / 32;
        dim3 gm((unsigned)((words + 255) / 256), (unsigned)C);
        if (p.detect_ts) { bsync_match_kernel<true><<<gm, 256, 0, h->stream>>>(p); }
        else { bsync_match_kernel<false><<<gm, 256, 0, h->stream>>>(p); }
        h->launches++;
    }
    cudaEventRecord(h->ev[2], h->stream);
    bsync_fsm_kernel<<<(C + 3) / 4, 128, 0, h->stream>>>(p);
    h->launches++;
    cudaEventRecord(h->ev[3], h->stream);
    BS_CUDA(cudaGetLastError());
    if (mem_kind == TDM_MEM_HOST) {
        if (n_bursts) { BS_CUDA(cudaMemcpyAsync(n_bursts, h->d_nbursts, sizeof(int) * (size_t)C, cudaMemcpyDeviceToHost, h->stream)); }
        if (bursts && max_bursts > 0) {
            BS_CUDA(cudaMemcpyAsync(bursts, h->d_bursts, sizeof(tdm_burst) * (size_t)C * (size_t)max_bursts, cudaMemcpyDeviceToHost, h->stream));
        }
        BS_CUDA(cudaStreamSynchronize(h->stream));
    }
    return TDM_OK;
}

int tdm_find_train_seq(int32_t device, void* cuda_stream, const uint8_t* in, int64_t in_stride, int32_t n_channels,
                       uint32_t end_of_in, uint32_t mask_of_train_seq, int32_t* out_type, uint32_t* out_offset,
                       int32_t mem_kind) {
    if (n_channels <= 0 || !out_type || !out_offset || (end_of_in > 0 && !in) || in_stride < (int64_t)end_of_in || end_of_in > (1u << 30)) {
        return tdm_internal_fail(TDM_ERR_ARG, "tdm_find_train_seq: bad arguments");
    }
    if (mem_kind != TDM_MEM_HOST && mem_kind != TDM_MEM_DEVICE) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_find_train_seq: mem_kind"); }
    int rc = check_device("tdm_find_train_seq", device);
    if (rc != TDM_OK) { return rc; }
    DevGuard guard(device);
    cudaStream_t stream = (cuda_stream == TDM_OWN_STREAM) ? nullptr : (cudaStream_t)cuda_stream;
    const size_t C = (size_t)n_channels;
    const long long wstride = words_for(end_of_in);
    uint32_t* d_wb = nullptr;
    uint8_t* d_in = nullptr;
    if (n_channels > 65535) { return tdm_internal_fail(TDM_ERR_UNSUPPORTED, "tdm_find_train_seq: at most 65535 channels per call (one grid row per channel)"); }
    int* d_type = nullptr;
    uint32_t* d_off = nullptr;
    auto done = [&](int code) { cudaFree(d_wb); cudaFree(d_in); cudaFree(d_type); cudaFree(d_off); return code; };
    if (cudaMalloc(&d_wb, sizeof(uint32_t) * C * (size_t)wstride) != cudaSuccess) { return done(tdm_internal_fail(TDM_ERR_NOMEM, "tdm_find_train_seq: cudaMalloc failed")); }
    BsyncParams p{};
    p.in = in; p.in_stride = in_stride; p.units_all = (int)end_of_in; p.max_units = (int)end_of_in; p.bits_per_unit = 1; p.n_channels = n_channels;
    p.wb = d_wb; p.wstride = wstride; p.find_end = end_of_in; p.find_mask = mask_of_train_seq;
    p.find_type = out_type; p.find_offset = out_offset;
    if (mem_kind == TDM_MEM_HOST) {
        if (cudaMalloc(&d_in, C * (size_t)(end_of_in ? end_of_in : 1)) != cudaSuccess || cudaMalloc(&d_type, sizeof(int) * C) != cudaSuccess ||
            cudaMalloc(&d_off, sizeof(uint32_t) * C) != cudaSuccess) {
            return done(tdm_internal_fail(TDM_ERR_NOMEM, "tdm_find_train_seq: cudaMalloc failed"));
        }
        if (end_of_in > 0 && cudaMemcpy2DAsync(d_in, end_of_in, in, (size_t)in_stride, end_of_in, C, cudaMemcpyHostToDevice, stream) != cudaSuccess) {
            return done(tdm_internal_fail(TDM_ERR_CUDA, "tdm_find_train_seq: copy in failed"));
        }
        p.in = d_in; p.in_stride = end_of_in; p.find_type = d_type; p.find_offset = d_off;
    }
    const long long row_bits = (long long)end_of_in + 32 * kPadWords;
    dim3 gpack((unsigned)((row_bits + kTileUnits - 1) / kTileUnits), (unsigned)n_channels);
    bsync_pack_kernel<<<gpack, 256, 0, stream>>>(p);
    bsync_find_kernel<<<(n_channels + 3) / 4, 128, 0, stream>>>(p);
    if (cudaGetLastError() != cudaSuccess) { return done(tdm_internal_fail(TDM_ERR_CUDA, "tdm_find_train_seq: launch failed")); }
    if (mem_kind == TDM_MEM_HOST) {
        cudaMemcpyAsync(out_type, d_type, sizeof(int) * C, cudaMemcpyDeviceToHost, stream);
        cudaMemcpyAsync(out_offset, d_off, sizeof(uint32_t) * C, cudaMemcpyDeviceToHost, stream);
    }
    if (cudaStreamSynchronize(stream) != cudaSuccess) { return done(tdm_internal_fail(TDM_ERR_CUDA, "tdm_find_train_seq: %s", cudaGetErrorString(cudaGetLastError()))); }
    return done(TDM_OK);
}

// tetra_burst_rx_cb's block split (phy/tetra_burst.c:33-49,343-393); host only.
int tdm_burst_unpack(const tdm_burst* b, uint8_t* bits510) {
    if (!b || !bits510) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_burst_unpack: null argument"); }
    for (int i = 0; i < TDM_BITS_PER_TS; ++i) { bits510[i] = (uint8_t)((b->bits[i >> 5] >> (31 - (i & 31))) & 1u); }
    return TDM_OK;
}

int tdm_burst_demux(const tdm_burst* b, tdm_tp_sap_block* out) {
    if (!b || !out) { return tdm_internal_fail(TDM_ERR_ARG, "tdm_burst_demux: null argument"); }
    enum { SB1 = 0, SB2 = 1, NDB = 2, BBK = 3, SCH_F = 5 };       // enum tp_sap_data_type, phy/tetra_burst.h:9-16
    uint8_t u[512];
    tdm_burst_unpack(b, u);
    std::memset(out, 0, 3 * sizeof(*out));
    auto put = [&](int k, int type, int blk, int off, int n, int at = 0) {
        out[k].type = type; out[k].blk_num = blk; out[k].n_bits = at + n;
        std::memcpy(out[k].bits + at, u + off, (size_t)n);
    };
    switch (b->train_seq) {
        case TDM_TRAIN_SYNC:                                        // SB1, broadcast block, SB2
            put(0, SB1, 1, (6 + 1 + 40) * 2, 120); put(1, BBK, 0, (6 + 1 + 40 + 60 + 19) * 2, 30); put(2, SB2, 2, (6 + 1 + 40 + 60 + 19 + 15) * 2, 216);
            return 3;
        case TDM_TRAIN_NORM_2:                                      // re-joined broadcast block, two separate blocks
            put(0, BBK, 0, (5 + 1 + 1 + 108) * 2, 14); put(0, BBK, 0, (5 + 1 + 1 + 108 + 7 + 11) * 2, 16, 14);
            put(1, NDB, 1, (5 + 1 + 1) * 2, 216); put(2, NDB, 2, (5 + 1 + 1 + 108 + 7 + 11 + 8) * 2, 216);
            return 3;
        case TDM_TRAIN_NORM_1:                                      // re-joined broadcast block, both blocks as one SCH/F
            put(0, BBK, 0, (5 + 1 + 1 + 108) * 2, 14); put(0, BBK, 0, (5 + 1 + 1 + 108 + 7 + 11) * 2, 16, 14);
            put(1, SCH_F, 0, (5 + 1 + 1) * 2, 216); put(1, SCH_F, 0, (5 + 1 + 1 + 108 + 7 + 11 + 8) * 2, 216, 216);
            return 2;
        default:
            return 0;
    }
}

}  // extern "C"
