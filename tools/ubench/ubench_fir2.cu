// FIR-role inner loop variants on B200 (development aid, not product): where do the taps come from and how are
// the samples loaded?  Each "call" = 7 ring blocks x 8 samples x 8 outputs = 448 FFMA2 (one far-role tick).
//   A: taps = constant bank (uniform operands), samples LDS.64            (round 1's role code)
//   B: taps = LDS.128 broadcast from shared memory into registers, samples LDS.128   (tdm_ws.cu fir_tick)
//   C: taps = constant bank, samples LDS.128
//   D: B, but every tap register is first moved through a uniform register? (not expressible) -> B with taps replicated as float2 pairs
#include <cstdio>
#include <cuda_runtime.h>
constexpr int T = 8, SLOTS = 16;
struct FP { float tpad[96]; };
struct Smem { float4 xs4[SLOTS * 4][32]; float2 out[T][32]; float tp[96]; float2 tp2[96]; };

__device__ __forceinline__ float2 fma2(float t, float2 h, float2 c) {
    float2 d;
    asm("{ .reg .b64 a, b, cc, dd;\n mov.b64 a, {%2, %2};\n mov.b64 b, {%3, %4};\n mov.b64 cc, {%5, %6};\n"
        " fma.rn.f32x2 dd, a, b, cc;\n mov.b64 {%0, %1}, dd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(t), "f"(h.x), "f"(h.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fma2v(float2 t, float2 h, float2 c) {
    float2 d;
    asm("{ .reg .b64 a, b, cc, dd;\n mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %5};\n mov.b64 cc, {%6, %7};\n"
        " fma.rn.f32x2 dd, a, b, cc;\n mov.b64 {%0, %1}, dd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(t.x), "f"(t.y), "f"(h.x), "f"(h.y), "f"(c.x), "f"(c.y));
    return d;
}

template <int MODE>
__device__ __forceinline__ void load(const FP& p, const Smem& sm, int lane, int qb, int s, float (&tt)[16], float2 (&tt2)[16], float2 (&h)[T]) {
    const int slot = (qb + s) & (SLOTS - 1);
    if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 15; ++c) { tt[c] = p.tpad[s * T + c]; }
        const float2* x2 = reinterpret_cast<const float2*>(&sm.xs4[slot * 4][0]);
#pragma unroll
        for (int j = 0; j < T; ++j) { h[j] = x2[(j * 32 + lane)]; }      // [8][32] float2 view: plain LDS.64
    } else {
        if (MODE == 1) {
            const float4* t4 = reinterpret_cast<const float4*>(sm.tp + 8 * s);
#pragma unroll
            for (int c = 0; c < 4; ++c) { const float4 v = t4[c]; tt[4 * c] = v.x; tt[4 * c + 1] = v.y; tt[4 * c + 2] = v.z; tt[4 * c + 3] = v.w; }
        } else if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 15; ++c) { tt[c] = p.tpad[s * T + c]; }
        } else {
            const float4* t4 = reinterpret_cast<const float4*>(sm.tp2 + 8 * s);
#pragma unroll
            for (int c = 0; c < 8; ++c) { const float4 v = t4[c]; tt2[2 * c] = make_float2(v.x, v.y); tt2[2 * c + 1] = make_float2(v.z, v.w); }
        }
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
            const float4 v = sm.xs4[slot * 4 + pr][lane];
            h[2 * pr] = make_float2(v.x, v.y); h[2 * pr + 1] = make_float2(v.z, v.w);
        }
    }
}
template <int MODE>
__device__ __forceinline__ void comp(const float (&tt)[16], const float2 (&tt2)[16], const float2 (&h)[T], float2 (&acc)[T]) {
#pragma unroll
    for (int j = 0; j < T; ++j) {
#pragma unroll
        for (int i = 0; i < T; ++i) {
            if (MODE == 3) { acc[i] = fma2v(tt2[j - i + 7], h[j], acc[i]); }
            else { acc[i] = fma2(tt[j - i + 7], h[j], acc[i]); }
        }
    }
}
template <int MODE, int WARPS, int STRUCT>
__global__ void __launch_bounds__(WARPS * 32) fir_bench(const __grid_constant__ FP p, long long* cyc, int reps, float2* sink) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *(reinterpret_cast<Smem*>(smem_raw) + (threadIdx.x >> 5));
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < SLOTS * 4; ++i) sm.xs4[i][lane] = make_float4(1.0f + i * 1e-3f, 2.0f - i * 1e-3f, 0.5f, 0.25f);
    for (int i = lane; i < 96; i += 32) { sm.tp[i] = p.tpad[i]; sm.tp2[i] = make_float2(p.tpad[i], p.tpad[i]); }
    __syncwarp();
    float2 tot = make_float2(0.f, 0.f);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        float2 acc[T];
#pragma unroll
        for (int i = 0; i < T; ++i) { acc[i] = make_float2(0.f, 0.f); }
        float ta[16], tb[16]; float2 ta2[16], tb2[16]; float2 ha[T], hb[T];
        load<MODE>(p, sm, lane, r, 0, ta, ta2, ha);
        if (STRUCT == 0) {
#pragma unroll 1
            for (int s = 0; s < 7; s += 2) {
                if (s + 1 < 7) { load<MODE>(p, sm, lane, r, s + 1, tb, tb2, hb); }
                comp<MODE>(ta, ta2, ha, acc);
                if (s + 2 < 7) { load<MODE>(p, sm, lane, r, s + 2, ta, ta2, ha); }
                if (s + 1 < 7) { comp<MODE>(tb, tb2, hb, acc); }
            }
        } else {
#pragma unroll 1
            for (int s = 0; s + 1 < 7; s += 2) {
                load<MODE>(p, sm, lane, r, s + 1, tb, tb2, hb);
                if (STRUCT == 2) { asm volatile("" ::: "memory"); }
                comp<MODE>(ta, ta2, ha, acc);
                load<MODE>(p, sm, lane, r, s + 2, ta, ta2, ha);
                if (STRUCT == 2) { asm volatile("" ::: "memory"); }
                comp<MODE>(tb, tb2, hb, acc);
            }
            comp<MODE>(ta, ta2, ha, acc);
        }
#pragma unroll
        for (int i = 0; i < T; ++i) sm.out[i][lane] = acc[i];
        tot.x += acc[0].x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    if (tot.x == 12345.f) sink[threadIdx.x] = tot;
}
template <int MODE, int WARPS, int STRUCT = 0> void run(const char* name) {
    FP p = {};
    for (int i = 0; i < 96; ++i) p.tpad[i] = 0.01f * i;
    long long* cyc; float2* sink; cudaMalloc(&cyc, 8); cudaMalloc(&sink, 1 << 16);
    const int reps = 2000;
    auto kk = fir_bench<MODE, WARPS, STRUCT>;
    cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Smem) * WARPS));
    kk<<<148, WARPS * 32, sizeof(Smem) * WARPS>>>(p, cyc, reps, sink);
    kk<<<148, WARPS * 32, sizeof(Smem) * WARPS>>>(p, cyc, reps, sink);
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-64s struct %d warps/CTA=%d: %7.1f cycles per 448 FFMA2 -> %.3f cycles per FFMA2 per warp (%s)\n", name, STRUCT, WARPS, (double)h / reps,
           (double)h / reps / 448, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0, 1, 0>("A taps const bank, samples LDS.64"); run<0, 1, 1>("A taps const bank, samples LDS.64"); run<0, 1, 2>("A taps const bank, samples LDS.64");
    run<1, 1, 0>("B taps LDS.128 -> regs, samples LDS.128"); run<1, 1, 1>("B taps LDS.128 -> regs, samples LDS.128"); run<1, 1, 2>("B taps LDS.128 -> regs, samples LDS.128");
    run<2, 1, 0>("C taps const bank, samples LDS.128"); run<2, 1, 1>("C taps const bank, samples LDS.128"); run<2, 1, 2>("C taps const bank, samples LDS.128");
    run<1, 4, 1>("B taps LDS.128 -> regs, samples LDS.128"); run<2, 4, 1>("C taps const bank, samples LDS.128"); run<0, 4, 1>("A taps const bank, samples LDS.64");
    run<1, 4, 2>("B taps LDS.128 -> regs, samples LDS.128"); run<2, 4, 2>("C taps const bank, samples LDS.128");
    return 0;
}
