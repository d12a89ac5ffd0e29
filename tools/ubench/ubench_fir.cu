// Isolated timing of the FIR role loop (development aid): one warp per SMSP, nothing else on the SM.
#include <cstdio>
#include "../../sdrpp_tetra_demodulator_b200/csrc/tdm_kernels.cu"
using namespace tdm;
struct FirSmem { float2 xs[kWsXEntries][32]; float2 out[8][32]; };
template <int NB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) fir_bench(const __grid_constant__ DemodParams p, long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FirSmem* sm = reinterpret_cast<FirSmem*>(smem_raw) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < kWsXEntries; ++i) sm->xs[i][lane] = make_float2(1.0f + i * 1e-3f, 2.0f - i * 1e-3f);
    __syncwarp();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        float acc[kWsT][2];
#pragma unroll
        for (int i = 0; i < kWsT; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
        ws_fir_blocks2<NB, 0>(p, sm->xs, lane, r, acc);
#pragma unroll
        for (int i = 0; i < kWsT; ++i) sm->out[i][lane] = make_float2(acc[i][0], acc[i][1]);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// heterogeneous: 4 warps, each a different instantiation (different code, like the real roles)
__global__ void __launch_bounds__(128) fir_bench_hetero(const __grid_constant__ DemodParams p, long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FirSmem* sm = reinterpret_cast<FirSmem*>(smem_raw) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = 0; i < kWsXEntries; ++i) sm->xs[i][lane] = make_float2(1.0f + i * 1e-3f, 2.0f - i * 1e-3f);
    __syncwarp();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        float acc[kWsT][2];
#pragma unroll
        for (int i = 0; i < kWsT; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
        if (w == 0) ws_fir_blocks2<7, 0>(p, sm->xs, lane, r, acc);
        else if (w == 1) ws_fir_blocks2<7, 1>(p, sm->xs, lane, r, acc);
        else if (w == 2) ws_fir_blocks2<9, 2>(p, sm->xs, lane, r, acc);
        else ws_fir_blocks2<8, 0>(p, sm->xs, lane, r, acc);
#pragma unroll
        for (int i = 0; i < kWsT; ++i) sm->out[i][lane] = make_float2(acc[i][0], acc[i][1]);
    }
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cyc[w] = t1 - t0;
}
void run_hetero() {
    DemodParams p = {};
    for (int f = 0; f < 3; ++f) for (int i = 0; i < TDM_TAP_PAD; ++i) p.tpad[f][i] = 0.01f * i;
    long long* cyc; cudaMalloc(&cyc, 64);
    const int reps = 2000;
    cudaFuncSetAttribute(fir_bench_hetero, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FirSmem) * 4));
    fir_bench_hetero<<<148, 128, sizeof(FirSmem) * 4>>>(p, cyc, reps);
    fir_bench_hetero<<<148, 128, sizeof(FirSmem) * 4>>>(p, cyc, reps);
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    const int nb[4] = {7, 7, 9, 8};
    for (int w = 0; w < 4; ++w) printf("hetero warp %d NB=%d: %.1f cycles per call -> %.3f cycles/FFMA\n", w, nb[w], (double)h[w] / reps, (double)h[w] / reps / (nb[w] * 128));
}
template <int NB, int WARPS> void run() {
    DemodParams p = {};
    for (int i = 0; i < TDM_TAP_PAD; ++i) p.tpad[0][i] = 0.01f * i;
    long long* cyc; cudaMalloc(&cyc, 8);
    const int reps = 2000;
    auto k = fir_bench<NB, WARPS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FirSmem) * WARPS));
    k<<<148, WARPS * 32, sizeof(FirSmem) * WARPS>>>(p, cyc, reps);
    k<<<148, WARPS * 32, sizeof(FirSmem) * WARPS>>>(p, cyc, reps);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("NB=%d warps/CTA=%d: %.1f cycles per call (%d FFMA) -> %.3f cycles/FFMA  err=%s\n", NB, WARPS, (double)h / reps,
           NB * 128, (double)h / reps / (NB * 128), cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<7, 1>(); run<7, 4>(); run<9, 4>(); run_hetero();
    return 0;
}
