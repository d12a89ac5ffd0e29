// What does an instruction of class X cost an SM sub-partition that is also streaming FFMA2?  (development aid)
// Three warps share every scheduler (warps w, w + 4, w + 8 of a 384-thread CTA): two run K FFMA2 each (8 independent
// chains; together they hold the FP32 pipe at its limit, 4 K cycles), the third runs K instructions of class X
// (8 independent chains).  Time of the slowest warp / K = 4 if X is free beside the FFMA2 streams, 4 + c if an X
// costs c cycles of whatever the scheduler is short of.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_mix ubench_mix.cu && ./ubench_mix
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
enum { kNone = 0, kFfma2, kFfma, kFmul, kFadd, kFmnmx, kIadd, kLop, kShf, kLds64, kLds128, kMufu, kSetpSel, kF2I, kFfma2Scalar, kFfmaImm, kFmulImm, kFaddImm, kSel, kImad, kSts64, kI2F, kMov, kN };
const char* kNames[kN] = { "(nothing)", "FFMA2", "FFMA 3-reg", "FMUL", "FADD", "FMNMX", "IADD3", "LOP3", "SHF", "LDS.64", "LDS.128", "MUFU.RSQ", "FSETP+FSEL", "F2I", "FFMA2 (scalar b)", "FFMA imm", "FMUL imm", "FADD imm", "SEL", "IMAD", "STS.64", "I2F", "MOV" };

template <int X, int PH>
__device__ __forceinline__ void body(unsigned long long (&v)[8], float (&f)[8], float (&g)[24], unsigned (&u)[8], unsigned long long b2, float fb, float fc, unsigned ub, unsigned smem_addr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (X == kFfma2) { asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(v[i]) : "l"(b2)); }
        if (X == kFfma2Scalar) { asm volatile("{ .reg .b64 t;\n mov.b64 t, {%1, %1};\n fma.rn.f32x2 %0, t, %0, %0; }" : "+l"(v[i]) : "f"(f[i])); }
        if (X == kFfma) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fb), "f"(fc)); }
        if (X == kFfmaImm) { asm volatile("fma.rn.f32 %0, %0, 0f3F7FF972, %1;" : "+f"(f[i]) : "f"(fc)); }
        if (X == kFmulImm) { asm volatile("mul.rn.f32 %0, %0, 0f3F7FF972;" : "+f"(f[i])); }
        if (X == kFaddImm) { asm volatile("add.rn.f32 %0, %0, 0f3F7FF972;" : "+f"(f[i])); }
        if (X == kSel) { asm volatile("slct.f32.s32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fc), "r"(ub)); }
        if (X == kImad) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(ub), "r"(ub + 1)); }
        if (X == kSts64) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(smem_addr + 256u * i), "f"(f[i]), "f"(fb) : "memory"); }
        if (X == kI2F) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(__float_as_uint(f[i]))); }
        if (X == kMov) { asm volatile("mov.b32 %0, %1;" : "=r"(u[i]) : "r"(u[(i + 1) & 7])); }
        if (X == kFmul) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb)); }
        if (X == kFadd) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fb)); }
        if (X == kFmnmx) { asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(PH ? fb : fc)); }
        if (X == kIadd) { asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(ub)); }
        if (X == kLop) { asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(ub)); }
        if (X == kShf) { asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(u[i]) : "r"(ub)); }
        if (X == kLds64) { asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f[i]), "=f"(g[i]) : "r"(smem_addr + 256u * i) : "memory"); }
        if (X == kLds128) { asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[i]), "=f"(g[i]), "=f"(g[8 + i]), "=f"(g[16 + i]) : "r"(smem_addr + 512u * i) : "memory"); }
        if (X == kMufu) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }
        if (X == kSetpSel) { asm volatile("{ .reg .pred p;\n setp.gt.f32 p, %0, %1;\n selp.f32 %0, %0, %2, p; }" : "+f"(f[i]) : "f"(fb), "f"(fc)); }
        if (X == kF2I) { asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i]))); }
    }
}

template <int X>
__device__ __forceinline__ long long run(float seed, unsigned smem_addr, float* sink) {
    unsigned long long v[8];
    float f[8], g[24];
    unsigned u[8];
    for (int i = 0; i < 24; ++i) { g[i] = seed; }
    for (int i = 0; i < 8; ++i) {
        f[i] = seed + i;
        u[i] = __float_as_uint(seed) + i;
        v[i] = ((unsigned long long)__float_as_uint(seed * 0.5f + i) << 32) | __float_as_uint(seed);
    }
    // per-thread values: uniform or immediate operands would select other instruction forms
    const unsigned long long b2 = ((unsigned long long)__float_as_uint(0.999f + 1e-6f * threadIdx.x) << 32) | __float_as_uint(1.001f - 1e-6f * threadIdx.x);
    const float fb = 0.9999f + 1e-7f * threadIdx.x, fc = 1e-3f * threadIdx.x;
    const unsigned ub = threadIdx.x * 7u + 1u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        body<X, 0>(v, f, g, u, b2, fb, fc, ub, smem_addr);
        body<X, 1>(v, f, g, u, b2, fb, fc, ub, smem_addr);
    }
    float acc = 0;
    for (int i = 0; i < 8; ++i) { acc += f[i] + g[i] + g[8 + i] + g[16 + i] + __uint_as_float(u[i]) + __uint_as_float((unsigned)v[i]) + __uint_as_float((unsigned)(v[i] >> 32)); }
    asm volatile("" :: "f"(acc) : "memory");          // the clock is read after every result exists
    const long long t1 = clock64();
    if (acc == 12345.678f) { *sink = acc; }
    return t1 - t0;
}

template <int XA, int XB>
__global__ void __launch_bounds__(384) k(float seed, long long* cyc, float* sink) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) { buf[i] = make_float4(seed, seed, seed, seed); }
    __syncthreads();
    const unsigned smem_addr = (unsigned)__cvta_generic_to_shared(buf) + 16u * (threadIdx.x & 31);
    const int warp = threadIdx.x >> 5;
    long long c;
    if (warp < 8) { c = run<XA>(seed, smem_addr, sink); }
    else { c = run<XB>(seed, smem_addr, sink); }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { cyc[warp] = c; }
}

template <int XA, int XB>
void go(long long* d_cyc, float* d_sink) {
    k<XA, XB><<<148, 384>>>(1.0f, d_cyc, d_sink);
    k<XA, XB><<<148, 384>>>(1.0f, d_cyc, d_sink);
    cudaDeviceSynchronize();
    long long h[12];
    cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = 16.0 * kIters;
    const long long m = h[0] > h[4] ? (h[0] > h[8] ? h[0] : h[8]) : (h[4] > h[8] ? h[4] : h[8]);
    printf("2 x %-10s + %-18s : warps finish after %.3f, %.3f, %.3f cycles per instruction of their own; all three: %.3f -> X costs %+.3f\n", kNames[XA], kNames[XB],
           h[0] / n, h[4] / n, h[8] / n, m / n, m / n - 4.0);
}

int main() {
    long long* d_cyc;
    float* d_sink;
    cudaMalloc(&d_cyc, 128);
    cudaMalloc(&d_sink, 4);
#define ROW(X) go<kFfma2, X>(d_cyc, d_sink); go<kNone, X>(d_cyc, d_sink);
    ROW(kNone) ROW(kFfma2) ROW(kFfma2Scalar) ROW(kFfma) ROW(kFfmaImm) ROW(kFmul) ROW(kFmulImm) ROW(kFadd) ROW(kFaddImm) ROW(kFmnmx) ROW(kSel) ROW(kSetpSel)
    ROW(kIadd) ROW(kLop) ROW(kShf) ROW(kImad) ROW(kF2I) ROW(kI2F) ROW(kMufu) ROW(kLds64) ROW(kLds128) ROW(kSts64)
    go<kFfma, kFfma>(d_cyc, d_sink);
    go<kFfma, kIadd>(d_cyc, d_sink);
    go<kIadd, kIadd>(d_cyc, d_sink);
    go<kFfma, kFfma2>(d_cyc, d_sink);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
