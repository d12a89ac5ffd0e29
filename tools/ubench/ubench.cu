// Micro-benchmarks for the B200 SM (development aid, not product): FFMA issue rate with register /
// constant-bank operands, dependent-issue latencies of the ops on the demodulator's serial chains.
#include <cstdio>
#include <cuda_runtime.h>
struct P { float t[64]; };
#define REP 4096
template <int MODE>
__global__ void k(const __grid_constant__ P p, float* out, long long* cyc, float seed) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
    float b = seed * 1.0001f, c = seed * 0.5f;
    long long t0 = clock64();
    if (MODE == 0) {          // 16 independent FFMA chains, register operands
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(a[i], b, c);
        }
    } else if (MODE == 1) {   // 16 independent chains, constant-bank multiplier
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(b, p.t[i], a[i]);
        }
    } else if (MODE == 2) {   // one dependent FFMA chain
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] = __fmaf_rn(a[0], b, c);
        }
    } else if (MODE == 3) {   // dependent FADD chain
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] = __fadd_rn(a[0], b);
        }
    } else if (MODE == 4) {   // dependent compare+select chain (FSETP -> FSEL)
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] = (a[0] > b) ? c : __fadd_rn(a[0], 1.0f);
        }
    } else if (MODE == 5) {   // dependent chain alternating FFMA and FSEL-type clamp (fminf/fmaxf)
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] = fminf(fmaxf(__fmaf_rn(a[0], b, c), -1e30f), 1e30f);
        }
    } else if (MODE == 6) {   // 4 independent chains
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i & 3] = __fmaf_rn(a[i & 3], b, c);
        }
    } else if (MODE == 7) {   // 8 independent chains
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i & 7] = __fmaf_rn(a[i & 7], b, c);
        }
    } else if (MODE == 8) {   // dependent sqrt chain (IEEE)
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] = __fsqrt_rn(__fadd_rn(a[0], 2.0f));
        }
    } else if (MODE == 9) {   // 16 independent chains, mixed FFMA (reg) + FADD
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) { a[i] = __fmaf_rn(a[i], b, c); a[i + 1] = __fadd_rn(a[i + 1], b); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, int ops_per_rep) {
    P p; for (int i = 0; i < 64; ++i) p.t[i] = 1.0f + i * 1e-3f;
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    k<MODE><<<148, threads>>>(p, out, cyc, 1.0f);
    k<MODE><<<148, threads>>>(p, out, cyc, 1.0f);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-46s threads/CTA %4d : %8.3f cycles per op per warp\n", name, threads, (double)h / ((double)REP * ops_per_rep));
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {32, 128, 256, 512}) {
        printf("--- %d threads per CTA (%d warps per SMSP)\n", th, th / 128 ? th / 128 : 1);
        if (th == 32) {
            run<2>("dependent FFMA chain (latency)", th, 16);
            run<3>("dependent FADD chain (latency)", th, 16);
            run<4>("dependent FSETP+FSEL+FADD (per iteration)", th, 16);
            run<5>("dependent FFMA+FMNMX+FMNMX (per iteration)", th, 16);
            run<8>("dependent FADD+sqrt_rn (per iteration)", th, 16);
            run<6>("4 independent FFMA chains", th, 16);
            run<7>("8 independent FFMA chains", th, 16);
        }
        run<0>("16 independent FFMA chains, reg operands", th, 16);
        run<1>("16 independent FFMA chains, const-bank operand", th, 16);
        run<9>("16 chains FFMA+FADD mixed", th, 16);
    }
    return 0;
}
