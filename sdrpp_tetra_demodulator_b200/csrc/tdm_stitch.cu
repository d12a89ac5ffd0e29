// tdm_stitch.cu -- kernels behind tdm_process_long (time-segment parallelism for ONE long
// capture).  The demodulation chain is a recurrence in time, so a single channel cannot be split exactly; but
// once its loops have converged, the decisions of a segment that started late from reset state are the same bits
// the sequential chain produces.  tdm_process_long therefore runs S overlapping segments of one capture as S
// "channels" of the batch kernel and joins their dibit streams here:
//
//   segment c covers samples [c L, (c+1) L + W): its first W samples are warm-up and are ALSO the last W samples
//   of segment c-1.  The join is found by content: the last K dibits of segment c-1 must appear exactly once in
//   segment c near symbol W/2; segment c contributes everything after that occurrence.  A segment whose warm-up
//   was not enough (no unique occurrence) is demodulated again, this time as the sequential continuation of its
//   predecessor (from the predecessor's final loop state, samples [c L + W, (c+1) L + W)), which needs no join.
//
// Bits only: the float loop states of later segments are not the sequential chain's (tdm_b200.h says so).
#include <cuda_runtime.h>
#include <stdint.h>
#include "tdm_kernels.cuh"

namespace tdm {
namespace {

// states[0] = the carried state of the logical channel, states[1..n) = a freshly initialised chain
__global__ void long_init_states_kernel(tdm_channel_state* states, const tdm_channel_state* carried, const tdm_channel_state* fresh, int n) {
    const int words = (int)(sizeof(tdm_channel_state) / 4);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * words; i += gridDim.x * blockDim.x) {
        const int c = i / words, k = i % words;
        reinterpret_cast<uint32_t*>(states + c)[k] = reinterpret_cast<const uint32_t*>(c == 0 ? carried : fresh)[k];
    }
}

// dst[c] = src[c - 1] for c >= 1: every segment becomes the continuation of its predecessor
__global__ void long_shift_states_kernel(tdm_channel_state* dst, const tdm_channel_state* src, int n) {
    const int words = (int)(sizeof(tdm_channel_state) / 4);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * words; i += gridDim.x * blockDim.x) {
        const int c = i / words, k = i % words;
        if (c >= 1) { reinterpret_cast<uint32_t*>(dst + c)[k] = reinterpret_cast<const uint32_t*>(src + c - 1)[k]; }
    }
}

// one warp per join c = 1 .. n-1
__global__ void stitch_find_kernel(const uint8_t* __restrict__ dib, long long stride, const int* __restrict__ counts, int n_rows,
                                   int K, int jlo, int jhi, int* __restrict__ join, const int* __restrict__ fixed) {
    const int lane = threadIdx.x & 31;
    const int c = 1 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_rows) { return; }
    if (fixed[c]) { return; }                                   // a redone segment starts where its predecessor ended: join 0
    const int cp = counts[c - 1], cc = counts[c];
    int found = -1, total = 0;
    if (cp >= K) {
        const uint8_t* __restrict__ tail = dib + (long long)(c - 1) * stride + (cp - K);
        const uint8_t* __restrict__ row = dib + (long long)c * stride;
        const int lo = jlo < K ? K : jlo, hi = jhi < cc ? jhi : cc;
        for (int j0 = lo; j0 <= hi; j0 += 32) {
            const int j = j0 + lane;
            bool ok = j <= hi;
            // compare backwards: the newest dibits disagree first when the candidate is wrong
            for (int k = 1; ok && k <= K; ++k) { ok = row[j - k] == tail[K - k]; }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (m) {
                if (found < 0) { found = j0 + __ffs(m) - 1; }
                total += __popc(m);
            }
        }
    }
    if (lane == 0) { join[c] = (total == 1) ? found : -1; }
}

__device__ __forceinline__ bool seg_resolved(const int* join, const int* fixed, int c) { return c == 0 || fixed[c] || join[c] >= 0; }

// how many segments still have no place in the stream; which of them can be redone now (predecessor settled)
// Which open segments can be redone now (predecessor settled), how many are left -- and which need no agreement at
// all: if the run that produced the predecessor's stream is not locked at the boundary (DQPSKSymbolExtractor::sync of
// its final state is down: a stretch without signal), there is nothing to agree on; such a segment is joined at the
// nominal place, symbol force_at of its stream.  force_all: give up on every open segment the same way (pass limit).
__global__ void stitch_plan_kernel(int* __restrict__ join, int* __restrict__ fixed, const int* __restrict__ counts, int n_rows,
                                   int* __restrict__ adopt, int* __restrict__ n_open, int* __restrict__ n_forced, int force_at, int force_all,
                                   const tdm_channel_state* __restrict__ final_states) {
    if (blockIdx.x != 0 || threadIdx.x != 0) { return; }
    int open = 0, forced = 0;
    for (int c = 1; c < n_rows; ++c) {
        if (!seg_resolved(join, fixed, c) && (force_all || final_states[c - 1].sync == 0u)) {
            join[c] = force_at < counts[c] ? force_at : counts[c];
            fixed[c] = 2;                                           // settled without agreement: not searched again
            ++forced;
        }
    }
    for (int c = 0; c < n_rows; ++c) {
        const bool un = !seg_resolved(join, fixed, c);
        adopt[c] = (un && seg_resolved(join, fixed, c - 1)) ? 1 : 0;
        open += un ? 1 : 0;
    }
    *n_open = open;
    *n_forced += forced;
}

// Is the redone stream of a segment (continuation of its predecessor) the same as the segment's own stream where both
// have had the whole segment to settle?  Both end at the same sample: compare their last K dibits, allowing for the
// symbol or two that can fall either side of the end.  One warp per segment that is about to be redone.
__global__ void stitch_verify_kernel(const uint8_t* __restrict__ dib, long long stride, const uint8_t* __restrict__ dib2, long long stride2,
                                     const int* __restrict__ counts, const int* __restrict__ counts2, const int* __restrict__ adopt, int n_rows,
                                     int K, int* __restrict__ agree) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_rows || !adopt[c]) { return; }
    const int n1 = counts[c], n2 = counts2[c];
    bool any = false;
    if (n1 >= K + 8 && n2 >= K + 8) {
        const uint8_t* __restrict__ a = dib + (long long)c * stride + (n1 - 4 - K);        // own stream, K dibits ending 4 before its end
        const uint8_t* __restrict__ b2 = dib2 + (long long)c * stride2;
        for (int d = -3; d <= 3 && !any; ++d) {
            const uint8_t* __restrict__ b = b2 + (n2 - 4 - K + d);
            bool ok = true;
            for (int k = lane; k < K; k += 32) { ok = ok && (a[k] == b[k]); }
            any = __all_sync(0xffffffffu, ok);
        }
    }
    if (lane == 0) { agree[c] = any ? 1 : 0; }
}

// What to do with each segment that was redone (one thread; reads the joins as they were BEFORE any take-over):
//   mode 1  take the redone stream (it agrees with the segment's own stream at the end, or there is no better witness);
//   mode 2  keep the segment's OWN stream, joined at the nominal place: the continuation of the predecessor disagrees
//           with it to the end, but the successor has joined the own stream by content -- two independent runs agree
//           with each other and not with the predecessor's, so it is the predecessor's state that is off (observed: a
//           run that started inside a stretch of noise can stay in a false lock for a million samples).
__global__ void stitch_decide_kernel(int* __restrict__ join, int* __restrict__ fixed, const int* __restrict__ counts, const int* __restrict__ adopt,
                                     const int* __restrict__ agree, int n_rows, int force_at, int* __restrict__ mode, int* __restrict__ n_forced) {
    if (blockIdx.x != 0 || threadIdx.x != 0) { return; }
    for (int c = 0; c < n_rows; ++c) {
        mode[c] = 0;
        if (!adopt[c]) { continue; }
        const bool successor_vouches = (c + 1 < n_rows) && !fixed[c + 1] && join[c + 1] >= 0;
        mode[c] = (agree[c] || !successor_vouches) ? 1 : 2;
    }
    for (int c = 0; c < n_rows; ++c) {
        if (mode[c] == 2) { join[c] = force_at < counts[c] ? force_at : counts[c]; fixed[c] = 2; *n_forced += 1; }
    }
}

// take over the redone segments: their stream replaces the segment's row, their final loop state becomes the segment's
__global__ void stitch_adopt_kernel(uint8_t* __restrict__ dib, long long stride, const uint8_t* __restrict__ dib2, long long stride2,
                                    int* __restrict__ counts, const int* __restrict__ counts2, int* __restrict__ join, int* __restrict__ fixed,
                                    const int* __restrict__ mode, tdm_channel_state* __restrict__ final_states,
                                    const tdm_channel_state* __restrict__ run_states) {
    const int c = blockIdx.y;
    if (mode[c] != 1) { return; }
    const int len = counts2[c];
    const uint8_t* __restrict__ src = dib2 + (long long)c * stride2;
    uint8_t* __restrict__ dst = dib + (long long)c * stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) { dst[i] = src[i]; }
    if (blockIdx.x == 0) {
        const int words = (int)(sizeof(tdm_channel_state) / 4);
        for (int k = threadIdx.x; k < words; k += blockDim.x) {
            reinterpret_cast<uint32_t*>(final_states + c)[k] = reinterpret_cast<const uint32_t*>(run_states + c)[k];
        }
    }
    __syncthreads();
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { counts[c] = len; join[c] = 0; fixed[c] = 1; }
}

// lengths and exclusive offsets of every segment's contribution (n <= a few thousand: one thread)
__global__ void stitch_scan_kernel(const int* __restrict__ counts, const int* __restrict__ join, int n_rows, long long* __restrict__ offs) {
    if (blockIdx.x != 0 || threadIdx.x != 0) { return; }
    long long acc = 0;
    for (int c = 0; c < n_rows; ++c) {
        offs[c] = acc;
        acc += (c == 0) ? counts[0] : (join[c] >= 0 ? counts[c] - join[c] : 0);
    }
    offs[n_rows] = acc;
}

__global__ void stitch_copy_kernel(const uint8_t* __restrict__ dib, long long stride, const int* __restrict__ counts,
                                   const int* __restrict__ join, const long long* __restrict__ offs, uint8_t* __restrict__ out, long long cap) {
    const int c = blockIdx.y;
    const int skip = (c == 0) ? 0 : join[c];
    if (skip < 0) { return; }
    const uint8_t* __restrict__ src = dib + (long long)c * stride + skip;
    const int len = counts[c] - skip;
    const long long o = offs[c];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        if (o + i < cap) { out[o + i] = src[i]; }
    }
}

// append n dibits of row `src` at out[offs[slot] ..) and advance offs[slot] (tail of the capture)
__global__ void stitch_append_kernel(const uint8_t* __restrict__ src, const int* __restrict__ count, long long* __restrict__ total,
                                     uint8_t* __restrict__ out, long long cap) {
    const long long o = *total;
    const int len = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        if (o + i < cap) { out[o + i] = src[i]; }
    }
}
__global__ void stitch_bump_kernel(const int* __restrict__ count, long long* __restrict__ total) { *total += *count; }

}  // namespace

void launch_long_init_states(tdm_channel_state* states, const tdm_channel_state* carried, const tdm_channel_state* fresh, int n, cudaStream_t s) {
    long_init_states_kernel<<<(n * 180 + 255) / 256, 256, 0, s>>>(states, carried, fresh, n);
}
void launch_long_shift_states(tdm_channel_state* dst, const tdm_channel_state* src, int n, cudaStream_t s) {
    long_shift_states_kernel<<<(n * 180 + 255) / 256, 256, 0, s>>>(dst, src, n);
}
void launch_stitch_find(const uint8_t* dib, long long stride, const int* counts, int n_rows, int K, int jlo, int jhi, int* join,
                        const int* fixed, cudaStream_t s) {
    if (n_rows < 2) { return; }
    stitch_find_kernel<<<(n_rows - 1 + 3) / 4, 128, 0, s>>>(dib, stride, counts, n_rows, K, jlo, jhi, join, fixed);
}
void launch_stitch_plan(int* join, int* fixed, const int* counts, int n_rows, int* adopt, int* n_open, int* n_forced, int force_at,
                        int force_all, const tdm_channel_state* final_states, cudaStream_t s) {
    stitch_plan_kernel<<<1, 32, 0, s>>>(join, fixed, counts, n_rows, adopt, n_open, n_forced, force_at, force_all, final_states);
}
void launch_stitch_adopt(uint8_t* dib, long long stride, const uint8_t* dib2, long long stride2, int* counts, const int* counts2, int* join,
                         int* fixed, const int* adopt, int* agree, int* mode, int* n_forced, int force_at, int K,
                         tdm_channel_state* final_states, const tdm_channel_state* run_states, int n_rows, long long max_len, cudaStream_t s) {
    long long gx = (max_len + 255) / 256;
    if (gx > 256) { gx = 256; }
    if (gx < 1) { gx = 1; }
    stitch_verify_kernel<<<(n_rows + 3) / 4, 128, 0, s>>>(dib, stride, dib2, stride2, counts, counts2, adopt, n_rows, K, agree);
    stitch_decide_kernel<<<1, 32, 0, s>>>(join, fixed, counts, adopt, agree, n_rows, force_at, mode, n_forced);
    stitch_adopt_kernel<<<dim3((unsigned)gx, (unsigned)n_rows), 256, 0, s>>>(dib, stride, dib2, stride2, counts, counts2, join, fixed, mode,
                                                                           final_states, run_states);
}
void launch_stitch_scan(const int* counts, const int* join, int n_rows, long long* offs, cudaStream_t s) {
    stitch_scan_kernel<<<1, 32, 0, s>>>(counts, join, n_rows, offs);
}
void launch_stitch_copy(const uint8_t* dib, long long stride, const int* counts, const int* join, const long long* offs, uint8_t* out,
                        long long cap, int n_rows, long long max_len, cudaStream_t s) {
    long long gx = (max_len + 255) / 256;
    if (gx > 512) { gx = 512; }
    if (gx < 1) { gx = 1; }
    stitch_copy_kernel<<<dim3((unsigned)gx, (unsigned)n_rows), 256, 0, s>>>(dib, stride, counts, join, offs, out, cap);
}
void launch_stitch_append(const uint8_t* src, const int* count, long long* total, uint8_t* out, long long cap, long long max_len, cudaStream_t s) {
    long long gx = (max_len + 255) / 256;
    if (gx > 512) { gx = 512; }
    if (gx < 1) { gx = 1; }
    stitch_append_kernel<<<(unsigned)gx, 256, 0, s>>>(src, count, total, out, cap);
    stitch_bump_kernel<<<1, 1, 0, s>>>(count, total);
}

}  // namespace tdm
