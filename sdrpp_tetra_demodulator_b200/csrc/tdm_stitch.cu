// tdm_stitch.cu -- kernels behind tdm_process_long / tdm_process_long_batch (time-segment
// parallelism for long captures of FEW channels).  The demodulation chain is a recurrence in time, so a channel
// cannot be split exactly; but once its loops have converged, the decisions of a run that started late from reset
// state are the same bits the sequential chain produces.  Each of C channels is therefore cut into S overlapping
// segments, all C*S of them run as rows of the batch kernel (row r = c*S + s), and their dibit streams are joined here:
//
//   segment s of a channel covers samples [s L, (s+1) L + W): its first W samples are warm-up and are ALSO the last W
//   samples of segment s-1.  The join is found by content: the last K dibits of segment s-1 must appear exactly once
//   in segment s near symbol W/2; segment s contributes everything after that occurrence.  A segment whose warm-up
//   was not enough (no unique occurrence) is demodulated again, this time as the sequential continuation of its
//   predecessor (from the final loop state of the run that produced the predecessor's stream, samples
//   [s L + W, (s+1) L + W)), which needs no join; its successor's join is then searched again against the new
//   stream, because two converged runs may disagree by one symbol about what lies before a segment boundary.
//
// Bits only: the float loop states of later segments are not the sequential chain's (tdm_b200.h says so).
// `S` below is segments per channel; rows with r % S == 0 are the first segment of a channel and never join.
#include <cuda_runtime.h>
#include <stdint.h>
#include "tdm_kernels.cuh"

namespace tdm {
namespace {

constexpr int kStateWords = (int)(sizeof(tdm_channel_state) / 4);

// states[r] = the carried state of its channel for first segments, a freshly initialised chain otherwise
__global__ void long_init_states_kernel(tdm_channel_state* states, const tdm_channel_state* carried, const tdm_channel_state* fresh, int n, int S) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * kStateWords; i += gridDim.x * blockDim.x) {
        const int r = i / kStateWords, k = i % kStateWords;
        const tdm_channel_state* src = (r % S == 0) ? carried + r / S : fresh;
        reinterpret_cast<uint32_t*>(states + r)[k] = reinterpret_cast<const uint32_t*>(src)[k];
    }
}

// dst[r] = src[r - 1] within a channel: every segment becomes the continuation of its predecessor
__global__ void long_shift_states_kernel(tdm_channel_state* dst, const tdm_channel_state* src, int n, int S) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * kStateWords; i += gridDim.x * blockDim.x) {
        const int r = i / kStateWords, k = i % kStateWords;
        if (r % S != 0) { reinterpret_cast<uint32_t*>(dst + r)[k] = reinterpret_cast<const uint32_t*>(src + r - 1)[k]; }
    }
}

// dst[c] = src[c * S + S - 1] (gather the last segment's state of every channel) or the reverse (scatter)
__global__ void long_last_states_kernel(tdm_channel_state* packed, tdm_channel_state* rows, int C, int S, int scatter) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * kStateWords; i += gridDim.x * blockDim.x) {
        const int c = i / kStateWords, k = i % kStateWords;
        uint32_t* a = reinterpret_cast<uint32_t*>(packed + c) + k;
        uint32_t* b = reinterpret_cast<uint32_t*>(rows + (long long)c * S + S - 1) + k;
        if (scatter) { *b = *a; } else { *a = *b; }
    }
}

// one warp per row; rows that start a channel or are settled without agreement are skipped
// `tails`: how long every row's stream is as a PREDECESSOR (its own length, or more once it has been extended past its
// end); `late` != 0: this is a retry against extended predecessors -- a success also records where the predecessor's
// contribution ends (cut) and settles the row for good (fixed = 3).
__global__ void stitch_find_kernel(const uint8_t* __restrict__ dib, long long stride, const int* __restrict__ counts, const int* __restrict__ tails,
                                   int n_rows, int S, int K, int jlo, int jhi, int* __restrict__ join, int* __restrict__ fixed,
                                   int* __restrict__ cut, int late) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_rows || c % S == 0) { return; }
    if (fixed[c]) { return; }                                   // settled: redone (join 0), joined late, or joined without agreement
    if (late && join[c] >= 0) { return; }
    const int cp = tails[c - 1], cc = counts[c];
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
    if (lane == 0) {
        const int j = (total == 1) ? found : -1;
        join[c] = j;
        if (late && j >= 0) { fixed[c] = 3; cut[c - 1] = cp; }
    }
}

__device__ __forceinline__ bool seg_resolved(const int* join, const int* fixed, int r, int S) { return r % S == 0 || fixed[r] || join[r] >= 0; }

// Which open segments can be redone now (predecessor settled), how many are left -- and which need no agreement at
// all: if the run that produced the predecessor's stream is not locked at the boundary (DQPSKSymbolExtractor::sync of
// its final state is down: a stretch without signal), there is nothing to agree on; such a segment is joined at the
// nominal place, symbol force_at of its stream.  force_all: give up on every open segment the same way (pass limit).
__global__ void stitch_plan_kernel(int* __restrict__ join, int* __restrict__ fixed, const int* __restrict__ counts, int n_rows, int S,
                                   int* __restrict__ adopt, int* __restrict__ n_open, int* __restrict__ n_forced, int force_at, int force_all,
                                   const tdm_channel_state* __restrict__ final_states) {
    if (blockIdx.x != 0 || threadIdx.x != 0) { return; }
    int open = 0, forced = 0;
    for (int r = 0; r < n_rows; ++r) {
        if (!seg_resolved(join, fixed, r, S) && (force_all || final_states[r - 1].sync == 0u)) {
            join[r] = force_at < counts[r] ? force_at : counts[r];
            fixed[r] = 2;                                           // settled without agreement: not searched again
            ++forced;
        }
    }
    for (int r = 0; r < n_rows; ++r) {
        const bool un = !seg_resolved(join, fixed, r, S);
        adopt[r] = (un && seg_resolved(join, fixed, r - 1, S)) ? 1 : 0;     // un implies r % S != 0, so r - 1 is the same channel
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
                                     const int* __restrict__ agree, int n_rows, int S, int force_at, int* __restrict__ mode, int* __restrict__ n_forced) {
    if (blockIdx.x != 0 || threadIdx.x != 0) { return; }
    for (int r = 0; r < n_rows; ++r) {
        mode[r] = 0;
        if (!adopt[r]) { continue; }
        const bool successor_vouches = (r + 1 < n_rows) && ((r + 1) % S != 0) && !fixed[r + 1] && join[r + 1] >= 0;
        mode[r] = (agree[r] || !successor_vouches) ? 1 : 2;
    }
    for (int r = 0; r < n_rows; ++r) {
        if (mode[r] == 2) { join[r] = force_at < counts[r] ? force_at : counts[r]; fixed[r] = 2; *n_forced += 1; }
    }
}

// take over the redone segments: their stream replaces the segment's row, their final loop state becomes the segment's
__global__ void stitch_adopt_kernel(uint8_t* __restrict__ dib, long long stride, const uint8_t* __restrict__ dib2, long long stride2,
                                    int* __restrict__ counts, const int* __restrict__ counts2, int* __restrict__ join, int* __restrict__ fixed,
                                    const int* __restrict__ mode, tdm_channel_state* __restrict__ final_states,
                                    const tdm_channel_state* __restrict__ run_states, int* __restrict__ cut, int n_rows) {
    for (int c = blockIdx.y; c < n_rows; c += gridDim.y) {             // grid.y is capped at 65535; c is uniform in a block
        if (mode[c] != 1) { continue; }
        const int len = counts2[c];
        const uint8_t* __restrict__ src = dib2 + (long long)c * stride2;
        uint8_t* __restrict__ dst = dib + (long long)c * stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) { dst[i] = src[i]; }
        if (blockIdx.x == 0) {
            for (int k = threadIdx.x; k < kStateWords; k += blockDim.x) {
                reinterpret_cast<uint32_t*>(final_states + c)[k] = reinterpret_cast<const uint32_t*>(run_states + c)[k];
            }
        }
        __syncthreads();
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
            counts[c] = len; join[c] = 0; fixed[c] = 1;
            // a successor that joined this row's EXTENDED old stream late has to look again: that stream is gone
            if (cut[c] > 0) {
                cut[c] = 0;
                if (c + 1 < n_rows && fixed[c + 1] == 3) { fixed[c + 1] = 0; join[c + 1] = -1; }
            }
        }
    }
}

// lengths of every segment's contribution and their exclusive offsets inside the channel's output row; per-channel totals
// (cut[r] > 0: row r was extended past its end and its late-joining successor takes over only at symbol cut[r])
__global__ void stitch_scan_kernel(const int* __restrict__ counts, const int* __restrict__ cut, const int* __restrict__ join, int n_rows, int S,
                                   long long* __restrict__ offs, long long* __restrict__ totals) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;                      // one thread per channel
    if (c * S >= n_rows) { return; }
    long long acc = 0;
    for (int s = 0; s < S; ++s) {
        const int r = c * S + s;
        offs[r] = acc;
        const int end = cut[r] > 0 ? cut[r] : counts[r];
        acc += (s == 0) ? end : (join[r] >= 0 ? end - join[r] : 0);
    }
    totals[c] = acc;
}

__global__ void stitch_copy_kernel(const uint8_t* __restrict__ dib, long long stride, const int* __restrict__ counts, const int* __restrict__ cut,
                                   const int* __restrict__ join, const long long* __restrict__ offs, int S, uint8_t* __restrict__ out, long long out_stride,
                                   int n_rows) {
    for (int r = blockIdx.y; r < n_rows; r += gridDim.y) {             // grid.y is capped at 65535
        const int skip = (r % S == 0) ? 0 : join[r];
        if (skip < 0) { continue; }
        const uint8_t* __restrict__ src = dib + (long long)r * stride + skip;
        const int len = (cut[r] > 0 ? cut[r] : counts[r]) - skip;
        const long long o = offs[r];
        uint8_t* __restrict__ dst = out + (long long)(r / S) * out_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
            if (o + i < out_stride) { dst[o + i] = src[i]; }
        }
    }
}

// append the tail run of every channel (row c of `src`, count[c] dibits) behind what the segments produced
__global__ void stitch_append_kernel(const uint8_t* __restrict__ src, long long stride, const int* __restrict__ count, long long* __restrict__ totals,
                                     uint8_t* __restrict__ out, long long out_stride, int C) {
    for (int c = blockIdx.y; c < C; c += gridDim.y) {
        const long long o = totals[c];
        const int len = count[c];
        const uint8_t* __restrict__ s = src + (long long)c * stride;
        uint8_t* __restrict__ dst = out + (long long)c * out_stride;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
            if (o + i < out_stride) { dst[o + i] = s[i]; }
        }
    }
}
__global__ void stitch_bump_kernel(const int* __restrict__ count, long long* __restrict__ totals, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { totals[c] += count[c]; }
}

unsigned blocks_for(long long n, int per, int cap) {
    long long g = (n + per - 1) / per;
    if (g > cap) { g = cap; }
    if (g < 1) { g = 1; }
    return (unsigned)g;
}

}  // namespace

void launch_long_init_states(tdm_channel_state* states, const tdm_channel_state* carried, const tdm_channel_state* fresh, int n, int S, cudaStream_t s) {
    long_init_states_kernel<<<blocks_for((long long)n * kStateWords, 256, 2048), 256, 0, s>>>(states, carried, fresh, n, S);
}
void launch_long_shift_states(tdm_channel_state* dst, const tdm_channel_state* src, int n, int S, cudaStream_t s) {
    long_shift_states_kernel<<<blocks_for((long long)n * kStateWords, 256, 2048), 256, 0, s>>>(dst, src, n, S);
}
void launch_long_last_states(tdm_channel_state* packed, tdm_channel_state* rows, int C, int S, int scatter, cudaStream_t s) {
    long_last_states_kernel<<<blocks_for((long long)C * kStateWords, 256, 2048), 256, 0, s>>>(packed, rows, C, S, scatter);
}
void launch_stitch_find(const uint8_t* dib, long long stride, const int* counts, const int* tails, int n_rows, int S, int K, int jlo, int jhi,
                        int* join, int* fixed, int* cut, int late, cudaStream_t s) {
    if (S < 2) { return; }
    stitch_find_kernel<<<(n_rows + 3) / 4, 128, 0, s>>>(dib, stride, counts, tails, n_rows, S, K, jlo, jhi, join, fixed, cut, late);
}
void launch_stitch_plan(int* join, int* fixed, const int* counts, int n_rows, int S, int* adopt, int* n_open, int* n_forced, int force_at,
                        int force_all, const tdm_channel_state* final_states, cudaStream_t s) {
    stitch_plan_kernel<<<1, 32, 0, s>>>(join, fixed, counts, n_rows, S, adopt, n_open, n_forced, force_at, force_all, final_states);
}
void launch_stitch_adopt(uint8_t* dib, long long stride, const uint8_t* dib2, long long stride2, int* counts, const int* counts2, int* join,
                         int* fixed, const int* adopt, int* agree, int* mode, int* n_forced, int force_at, int K,
                         tdm_channel_state* final_states, const tdm_channel_state* run_states, int* cut, int n_rows, int S, long long max_len,
                         cudaStream_t s) {
    stitch_verify_kernel<<<(n_rows + 3) / 4, 128, 0, s>>>(dib, stride, dib2, stride2, counts, counts2, adopt, n_rows, K, agree);
    stitch_decide_kernel<<<1, 32, 0, s>>>(join, fixed, counts, adopt, agree, n_rows, S, force_at, mode, n_forced);
    stitch_adopt_kernel<<<dim3(blocks_for(max_len, 256, 256), (unsigned)(n_rows < 65535 ? n_rows : 65535)), 256, 0, s>>>(dib, stride, dib2, stride2, counts, counts2, join, fixed,
                                                                                              mode, final_states, run_states, cut, n_rows);
}
void launch_stitch_scan(const int* counts, const int* cut, const int* join, int n_rows, int S, long long* offs, long long* totals, cudaStream_t s) {
    const int C = n_rows / S;
    stitch_scan_kernel<<<(C + 127) / 128, 128, 0, s>>>(counts, cut, join, n_rows, S, offs, totals);
}
void launch_stitch_copy(const uint8_t* dib, long long stride, const int* counts, const int* cut, const int* join, const long long* offs, int S,
                        uint8_t* out, long long out_stride, int n_rows, long long max_len, cudaStream_t s) {
    stitch_copy_kernel<<<dim3(blocks_for(max_len, 256, 512), (unsigned)(n_rows < 65535 ? n_rows : 65535)), 256, 0, s>>>(dib, stride, counts, cut, join, offs, S, out, out_stride, n_rows);
}
void launch_stitch_append(const uint8_t* src, long long stride, const int* count, long long* totals, uint8_t* out, long long out_stride, int C,
                          long long max_len, cudaStream_t s) {
    stitch_append_kernel<<<dim3(blocks_for(max_len, 256, 64), (unsigned)(C < 65535 ? C : 65535)), 256, 0, s>>>(src, stride, count, totals, out, out_stride, C);
    stitch_bump_kernel<<<(C + 127) / 128, 128, 0, s>>>(count, totals, C);
}

}  // namespace tdm
