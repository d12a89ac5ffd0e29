// FFMA2 (fma.rn.f32x2) on B200: issue rate, latency, bit-exactness against two scalar fmas, and the
// FIR-role inner loop built from it (development aid, not product).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
struct P { float t[96]; };
#define REP 2048

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

// MODE 0: 16 independent FFMA2 chains, all-register vector operands
// MODE 1: 16 independent FFMA2 chains, scalar multiplier from the constant bank (uniform register broadcast)
// MODE 2: 16 independent FFMA2 chains, scalar multiplier in a per-thread register
// MODE 3: one dependent FFMA2 chain (latency)
// MODE 4: 32 independent scalar FFMA chains, const-bank multiplier (same flops as MODE 1)
// MODE 5: 8 independent FFMA2 chains (const-bank scalar)
template <int MODE>
__global__ void k(const __grid_constant__ P p, float2* out, long long* cyc, float seed, int active_mask) {
    float2 a[16];
    for (int i = 0; i < 16; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    float2 b = make_float2(seed * 1.0001f, seed * 0.9999f), c = make_float2(seed * 0.5f, seed * 0.25f);
    float bs = seed * 1.0001f + threadIdx.x * 1e-6f;
    const int warp = threadIdx.x >> 5;
    if (!((active_mask >> warp) & 1)) { return; }
    long long t0 = clock64();
    if (MODE == 0) {
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2v(b, c, a[i]);
        }
    } else if (MODE == 1) {
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(p.t[i], b, a[i]);
        }
    } else if (MODE == 2) {
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(bs, b, a[i]);
        }
    } else if (MODE == 3) {
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[0] = fma2v(a[0], b, c);
        }
    } else if (MODE == 4) {
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i].x = __fmaf_rn(p.t[i], b.x, a[i].x); a[i].y = __fmaf_rn(p.t[i], b.y, a[i].y); }
        }
    } else if (MODE == 5) {
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i & 7] = fma2(p.t[i], b, a[i & 7]);
        }
    }
    long long t1 = clock64();
    float2 s = make_float2(0.f, 0.f);
    for (int i = 0; i < 16; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
}
template <int MODE> void run(const char* name, int threads, int mask, int ops_per_rep) {
    P p; for (int i = 0; i < 96; ++i) p.t[i] = 1.0f + i * 1e-3f;
    float2* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 256); cudaMemset(cyc, 0, 256);
    k<MODE><<<148, threads>>>(p, out, cyc, 1.0f, mask);
    k<MODE><<<148, threads>>>(p, out, cyc, 1.0f, mask);
    long long h[32]; cudaMemcpy(h, cyc, 256, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int w = 0; w < threads / 32; ++w) if ((mask >> w) & 1) mx = h[w] > mx ? h[w] : mx;
    printf("%-58s threads %4d mask 0x%02x : %7.3f cycles per instr per warp (%s)\n", name, threads, mask,
           (double)mx / ((double)REP * ops_per_rep), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

// ---- bit-exactness: FFMA2 halves vs __fmaf_rn
__global__ void exact_k(const float* a, const float* b, const float* c, int n, unsigned* mism) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 >= n) return;
    float2 r = fma2v(make_float2(a[2 * i], a[2 * i + 1]), make_float2(b[2 * i], b[2 * i + 1]), make_float2(c[2 * i], c[2 * i + 1]));
    float2 s = fma2(a[2 * i], make_float2(b[2 * i], b[2 * i + 1]), make_float2(c[2 * i], c[2 * i + 1]));
    float e0 = __fmaf_rn(a[2 * i], b[2 * i], c[2 * i]), e1 = __fmaf_rn(a[2 * i + 1], b[2 * i + 1], c[2 * i + 1]);
    float e2 = __fmaf_rn(a[2 * i], b[2 * i + 1], c[2 * i + 1]);
    if (__float_as_uint(r.x) != __float_as_uint(e0) || __float_as_uint(r.y) != __float_as_uint(e1) ||
        __float_as_uint(s.x) != __float_as_uint(e0) || __float_as_uint(s.y) != __float_as_uint(e2)) atomicAdd(mism, 1u);
}
void run_exact() {
    const int n = 1 << 22;
    float *h = (float*)malloc(3 * n * sizeof(float));
    srand(1);
    for (int i = 0; i < 3 * n; ++i) {
        unsigned u = ((unsigned)rand() << 16) ^ (unsigned)rand() ^ ((unsigned)rand() << 31);
        int cls = rand() % 16;
        if (cls == 0) u &= 0x807fffffu;                                   // denormal / zero
        else if (cls < 12) u = (u & 0x807fffffu) | ((unsigned)(120 + rand() % 16) << 23);   // near 1
        float f; memcpy(&f, &u, 4);
        if (std::isnan(f) ) f = 1.0f;
        h[i] = f;
    }
    float* d; unsigned* m; cudaMalloc(&d, 3 * n * sizeof(float)); cudaMalloc(&m, 4); cudaMemset(m, 0, 4);
    cudaMemcpy(d, h, 3 * n * sizeof(float), cudaMemcpyHostToDevice);
    exact_k<<<(n / 2 + 255) / 256, 256>>>(d, d + n, d + 2 * n, n, m);
    unsigned mm; cudaMemcpy(&mm, m, 4, cudaMemcpyDeviceToHost);
    printf("FFMA2 vs __fmaf_rn on %d random triples (incl. denormals, inf): %u mismatches (%s)\n", n / 2, mm, cudaGetErrorString(cudaGetLastError()));
}

// ---- FIR role inner loop with FFMA2: T x T blocks, taps from the constant bank, samples from a shared ring
constexpr int T = 8, SLOTS = 16;
struct FirSmem { float2 xs[SLOTS * T][32]; float2 out[T][32]; };
struct FP { float tpad[3][88]; };
template <int NB, int F>
__device__ __forceinline__ void fir_blocks_f2(const FP& p, const float2 (*xs)[32], int lane, int qblock0, float2 (&acc)[T]) {
    float ta[2 * T - 1], tb[2 * T - 1];
    float2 ha[T], hb[T];
    auto load = [&](float (&tt)[2 * T - 1], float2 (&h)[T], int s) {
#pragma unroll
        for (int c = 0; c < 2 * T - 1; ++c) { tt[c] = p.tpad[F][s * T + c]; }
        const int slot = (qblock0 + s) & (SLOTS - 1);
#pragma unroll
        for (int j = 0; j < T; ++j) { h[j] = xs[slot * T + j][lane]; }
    };
    auto comp = [&](const float (&tt)[2 * T - 1], const float2 (&h)[T]) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
#pragma unroll
            for (int i = 0; i < T; ++i) { acc[i] = fma2(tt[j - i + T - 1], h[j], acc[i]); }
        }
    };
    load(ta, ha, 0);
#pragma unroll 1
    for (int s = 0; s + 1 < NB; s += 2) {
        load(tb, hb, s + 1);
        comp(ta, ha);
        load(ta, ha, s + 2);
        comp(tb, hb);
    }
    if (NB & 1) { comp(ta, ha); }
}
// single-buffered, fully unrolled over the blocks: taps become immediate constant-bank operands
template <int NB, int F>
__device__ __forceinline__ void fir_blocks_f2u(const FP& p, const float2 (*xs)[32], int lane, int qblock0, float2 (&acc)[T]) {
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        const int slot = (qblock0 + s) & (SLOTS - 1);
        float2 h[T];
#pragma unroll
        for (int j = 0; j < T; ++j) { h[j] = xs[slot * T + j][lane]; }
#pragma unroll
        for (int j = 0; j < T; ++j) {
#pragma unroll
            for (int i = 0; i < T; ++i) { acc[i] = fma2(p.tpad[F][s * T + j - i + T - 1], h[j], acc[i]); }
        }
    }
}
template <int NB, int WARPS, int UNROLLED>
__global__ void __launch_bounds__(WARPS * 32) fir_bench(const __grid_constant__ FP p, long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FirSmem* sm = reinterpret_cast<FirSmem*>(smem_raw) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < SLOTS * T; ++i) sm->xs[i][lane] = make_float2(1.0f + i * 1e-3f, 2.0f - i * 1e-3f);
    __syncwarp();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        float2 acc[T];
#pragma unroll
        for (int i = 0; i < T; ++i) { acc[i] = make_float2(0.f, 0.f); }
        if (UNROLLED) fir_blocks_f2u<NB, 0>(p, sm->xs, lane, r, acc);
        else fir_blocks_f2<NB, 0>(p, sm->xs, lane, r, acc);
#pragma unroll
        for (int i = 0; i < T; ++i) sm->out[i][lane] = acc[i];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NB, int WARPS, int UNROLLED> void run_fir() {
    FP p = {};
    for (int f = 0; f < 3; ++f) for (int i = 0; i < 88; ++i) p.tpad[f][i] = 0.01f * i;
    long long* cyc; cudaMalloc(&cyc, 8);
    const int reps = 2000;
    auto kk = fir_bench<NB, WARPS, UNROLLED>;
    cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FirSmem) * WARPS));
    kk<<<148, WARPS * 32, sizeof(FirSmem) * WARPS>>>(p, cyc, reps);
    kk<<<148, WARPS * 32, sizeof(FirSmem) * WARPS>>>(p, cyc, reps);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("FIR f32x2 %s NB=%d warps/CTA=%d: %.1f cycles per call (%d FMA) -> %.3f cycles/FMA  (%s)\n", UNROLLED ? "unrolled" : "rolled  ",
           NB, WARPS, (double)h / reps, NB * 128, (double)h / reps / (NB * 128), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run_exact();
    run<3>("dependent FFMA2 chain (latency)", 32, 1, 16);
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int threads = cfg == 0 ? 32 : cfg == 1 ? 128 : 256;
        const int mask = cfg == 0 ? 1 : cfg == 1 ? 0xf : 0x11;      // 1 warp; 1 warp per SMSP; 2 warps on one SMSP
        run<0>("16 indep FFMA2, vector regs", threads, mask, 16);
        run<1>("16 indep FFMA2, const-bank scalar (UR.F32)", threads, mask, 16);
        run<2>("16 indep FFMA2, register scalar", threads, mask, 16);
        run<5>("8 indep FFMA2, const-bank scalar", threads, mask, 16);
        run<4>("32 indep scalar FFMA, const-bank (2 per 'instr')", threads, mask, 16);
    }
    run_fir<7, 1, 0>(); run_fir<7, 4, 0>(); run_fir<7, 8, 0>(); run_fir<9, 4, 0>();
    run_fir<7, 1, 1>(); run_fir<7, 4, 1>(); run_fir<7, 8, 1>(); run_fir<9, 4, 1>();
    return 0;
}
