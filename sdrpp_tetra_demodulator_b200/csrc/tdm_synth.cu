// tdm_synth.cu -- deterministic synthetic TETRA-mapped pi/4-DQPSK captures, generated
// directly in HBM (test / benchmark signal source; the reference ships no captures,
// SURVEY.md 4 and 8d).  The symbol mapping is the reference's bits2phase table
// (src/decoder/src/phy/tetra_burst.c:99-104): 00 -> +pi/4, 01 -> +3pi/4, 11 -> -3pi/4,
// 10 -> -pi/4.
//
// Recipe (integer parts are exactly reproducible on any host, see tests/):
//   a_k     = hash(seed_data+c, 16+k) & 3         idx_k = 2 a_k + (k & 1)     (units of pi/4)
//   dibit_k = map[(idx_k - idx_{k-1}) & 7],  map: 1->00, 3->01, 5->11, 7->10,  idx_{-1} = 7
//   s[n]    = A e^{j(2 pi df n/fs + phi0)} sum_k e^{j pi idx_k/4} h(n - 2k - 2 tau) + w[n]
//   h       = RRC beta 0.35, Ts = 2 samples, support |t| <= 33;  w = AWGN at Es/N0 = snr_db
//   df, tau, A, phi0 = draws 0..3 of hash(seed_data+c, .);  noise from hash(seed_noise+c, n)
#include "tdm_kernels.cuh"

namespace tdm {
namespace {

constexpr int kSpan = 33;
constexpr int kTab = 2 * kSpan + 2;
constexpr int kBlockSamples = 2048;
constexpr double kPi = 3.14159265358979323846;

__host__ __device__ inline unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ inline unsigned long long hash2(unsigned long long seed, unsigned long long idx) {
    return mix64(mix64(seed * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL) ^ (idx * 0x9E3779B97F4A7C15ULL));
}
__host__ __device__ inline double u01(unsigned long long h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
__host__ __device__ inline int abs_index(unsigned long long seed, long long k) {
    if (k < 0) { return 7; } /* virtual symbol -1: keeps the first increment odd */
    return (int)(2 * (hash2(seed, 16 + (unsigned long long)k) & 3) + (unsigned long long)(k & 1));
}

__device__ double rrc_pulse(double t, double beta) {
    const double Ts = 2.0, x = t / Ts;
    if (fabs(t) < 1e-12) { return 1.0 + beta * (4.0 / kPi - 1.0); }
    if (fabs(fabs(4.0 * beta * x) - 1.0) < 1e-9) {
        return (beta / sqrt(2.0)) * ((1.0 + 2.0 / kPi) * sin(kPi / (4.0 * beta)) + (1.0 - 2.0 / kPi) * cos(kPi / (4.0 * beta)));
    }
    const double num = sin(kPi * x * (1.0 - beta)) + 4.0 * beta * x * cos(kPi * x * (1.0 + beta));
    const double den = kPi * x * (1.0 - (4.0 * beta * x) * (4.0 * beta * x));
    return num / den;
}

__global__ void __launch_bounds__(256) synth_kernel(tdm_synth_params sp, int first_channel, long long n_samples,
                                                     long long stride, float2* __restrict__ iq,
                                                     uint8_t* __restrict__ tx, long long tx_stride) {
    __shared__ float tab[kTab];
    __shared__ unsigned char sidx[kBlockSamples / 2 + kSpan + 4];
    __shared__ float s_sigma, s_amp;
    __shared__ unsigned long long s_inc, s_ph0;
    const int cl = blockIdx.y;                 // channel within this call's buffer
    const int c = first_channel + cl;          // global channel number -> seeds
    const unsigned long long sd = sp.seed_data + (unsigned long long)c, sn = sp.seed_noise + (unsigned long long)c;
    const long long n0 = (long long)blockIdx.x * kBlockSamples;
    if (n0 >= n_samples) { return; }

    const double tau = u01(hash2(sd, 1));
    if (threadIdx.x < kTab) {
        const int j = (int)threadIdx.x - kSpan;
        const double t = (double)j - 2.0 * tau;
        tab[threadIdx.x] = (fabs(t) <= (double)kSpan) ? (float)rrc_pulse(t, 0.35) : 0.f;
    }
    // symbol indices needed by this block: j = n - 2k in [-kSpan, kSpan+1]
    const long long kfirst = (n0 - (kSpan + 1)) / 2 - 1;
    for (int i = threadIdx.x; i < kBlockSamples / 2 + kSpan + 4; i += blockDim.x) {
        const long long k = kfirst + i;
        sidx[i] = (k >= 0) ? (unsigned char)abs_index(sd, k) : (unsigned char)255;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double e = 0.0;
        for (int j = 0; j < kTab; ++j) { e += (double)tab[j] * (double)tab[j]; }
        const double amp = sp.min_amp * pow(sp.max_amp / sp.min_amp, u01(hash2(sd, 2)));
        const double ps = amp * amp * e / 2.0;
        const double sigma2 = ps * 2.0 / pow(10.0, sp.snr_db / 10.0);
        s_sigma = (float)sqrt(sigma2 / 2.0);
        s_amp = (float)amp;
        const double df = (2.0 * u01(hash2(sd, 0)) - 1.0) * sp.max_freq_off_hz;
        s_inc = (unsigned long long)(long long)llrint(df / 36000.0 * 18446744073709551616.0);
        const double phi0 = 2.0 * kPi * u01(hash2(sd, 3));
        s_ph0 = ((unsigned long long)llrint(phi0 / (2.0 * kPi) * 4294967296.0)) << 32;
    }
    __syncthreads();
    const float sigma = s_sigma, amp = s_amp;
    const float cs8x[8] = { 1.f, 0.70710678f, 0.f, -0.70710678f, -1.f, -0.70710678f, 0.f, 0.70710678f };
    float2* row = iq + (long long)cl * stride;

    for (int i = threadIdx.x; iq != nullptr && i < kBlockSamples; i += blockDim.x) {
        const long long n = n0 + i;
        if (n >= n_samples) { break; }
        long long kmin = (n - (kSpan + 1) + 1) / 2;
        if (n - (kSpan + 1) < 0) { kmin = 0; }
        const long long kmax = (n + kSpan) / 2;
        float re = 0.f, im = 0.f;
        for (long long k = kmin; k <= kmax; ++k) {
            const int j = (int)(n - 2 * k);
            if (j < -kSpan || j > kSpan + 1) { continue; }
            const int idx = sidx[(int)(k - kfirst)];
            const float h = tab[j + kSpan];
            re = fmaf(cs8x[idx & 7], h, re);
            im = fmaf(cs8x[(idx + 6) & 7], h, im);
        }
        const unsigned long long ph = s_ph0 + s_inc * (unsigned long long)n;
        const float turns = (float)(unsigned int)(ph >> 32) * (1.0f / 4294967296.0f);
        float sr, cr;
        sincospif(2.0f * turns, &sr, &cr);
        cr *= amp; sr *= amp;
        const float vr = re * cr - im * sr, vi = re * sr + im * cr;
        const unsigned long long hn = hash2(sn, (unsigned long long)n);
        const float u1 = ((float)(unsigned int)(hn >> 40) + 0.5f) * (1.0f / 16777216.0f);
        const float u2 = ((float)(unsigned int)(hn & 0xFFFFFFu) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sigma * sqrtf(-2.0f * __logf(u1));
        float s2, c2;
        sincospif(2.0f * u2, &s2, &c2);
        row[n] = make_float2(vr + r * c2, vi + r * s2);
    }

    // transmitted dibits for the symbols whose first sample falls in this block
    if (tx) {
        const int map8[8] = { 0, 0, 0, 1, 0, 3, 0, 2 };
        for (int i = threadIdx.x; i < kBlockSamples / 2; i += blockDim.x) {
            const long long k = n0 / 2 + i;
            if (2 * k >= n_samples || k >= tx_stride) { break; }
            const int d = (abs_index(sd, k) - abs_index(sd, k - 1)) & 7;
            tx[(long long)cl * tx_stride + k] = (uint8_t)map8[d];
        }
    }
}

}  // namespace

int launch_synth(const tdm_synth_params& sp, int n_channels, long long n_samples, long long stride,
                 int first_channel, float2* iq, uint8_t* tx_dibits, long long tx_stride, cudaStream_t stream) {
    if (n_channels <= 0 || n_samples <= 0) { return 0; }
    int launches = 0;
    const long long nblk = (n_samples + kBlockSamples - 1) / kBlockSamples;
    // grid.y is limited to 65535 channels per launch
    for (int c0 = 0; c0 < n_channels; c0 += 32768) {
        const int nc = (n_channels - c0) < 32768 ? (n_channels - c0) : 32768;
        dim3 grid((unsigned)nblk, (unsigned)nc);
        synth_kernel<<<grid, 256, 0, stream>>>(sp, first_channel + c0, n_samples, stride, iq ? iq + (long long)c0 * stride : nullptr,
                                               tx_dibits ? tx_dibits + (long long)c0 * tx_stride : nullptr, tx_stride);
        if (cudaGetLastError() != cudaSuccess) { return -1; }
        ++launches;
    }
    return launches;
}

}  // namespace tdm
