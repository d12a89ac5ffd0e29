/* A wideband receiver without Python: one capture of interleaved int16 IQ (what SDR hardware and SDR++'s baseband recorder
 * deliver) -> front-end channeliser (include/tdm_chan_b200.h: every channel of the 25 kHz raster at 36 kS/s, where the
 * reference asks SDR++ for one VFO per plugin instance, src/main.cpp:75) -> demodulator (include/tdm_b200.h), both on the
 * GPU, the channel samples handed over instant-major and never leaving HBM.  Prints the channels whose lock metric says
 * "synchronised" (DQPSKSymbolExtractor::sync, src/dsp/dqpsk_sym_extr.cpp:25) and how many symbols each produced.
 *
 *   gcc -std=c11 -I include -I /usr/local/cuda/include examples/wideband_chain.c \
 *       -L sdrpp_tetra_demodulator_b200 -ltdm_b200 -L /usr/local/cuda/lib64 -lcudart -o wideband_chain
 *   ./wideband_chain capture.cs16 G          (fs = 0.9 G MHz, 36 G channels; G = 4: 3.6 MS/s, 144 channels)
 *
 * Without a B200 it stops at tdm_chan_create with TDM_ERR_NO_DEVICE: there is no CPU fallback. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "tdm_b200.h"
#include "tdm_chan_b200.h"

#define CHECK_TDM(call) do { int rc_ = (call); if (rc_ != TDM_OK) { fprintf(stderr, "%s: status %d: %s\n", #call, rc_, tdm_last_error()); return rc_ == TDM_ERR_NO_DEVICE ? 3 : 1; } } while (0)
#define CHECK_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: wideband_chain capture.cs16 G\n"); return 2; }
    const int32_t g = atoi(argv[2]);
    tdm_chan_config cc;
    CHECK_TDM(tdm_chan_default_config(g, &cc));
    const int32_t M = cc.n_channels, D = cc.decimation;
    const int32_t instants = 4096;                               /* channel samples per call */
    tdm_chan* chan = NULL;
    CHECK_TDM(tdm_chan_create(&cc, 0, &chan));                   /* no B200: TDM_ERR_NO_DEVICE */
    tdm_config cfg;
    tdm_default_config(&cfg);
    tdm_handle* demod = NULL;
    CHECK_TDM(tdm_create(&cfg, M, instants, 0, &demod));
    CHECK_TDM(tdm_set_stream(demod, NULL));                      /* one stream (the default one) for both stages */

    FILE* f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    const size_t n_wide = (size_t)instants * (size_t)D;
    const int64_t S = tdm_max_symbols(demod, instants);
    int16_t* h_wide = NULL;
    int16_t* d_wide; float* d_inst; uint8_t* d_dibits; int32_t* d_counts;
    CHECK_CUDA(cudaMallocHost((void**)&h_wide, n_wide * 2 * sizeof(int16_t)));
    CHECK_CUDA(cudaMalloc((void**)&d_wide, n_wide * 2 * sizeof(int16_t)));
    CHECK_CUDA(cudaMalloc((void**)&d_inst, (size_t)instants * (size_t)M * 2 * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&d_dibits, (size_t)M * (size_t)S));
    CHECK_CUDA(cudaMalloc((void**)&d_counts, (size_t)M * sizeof(int32_t)));

    size_t got;
    long long wide_total = 0;
    while ((got = fread(h_wide, 2 * sizeof(int16_t), n_wide, f)) >= (size_t)D) {
        const int64_t n = (int64_t)(got / (size_t)D) * D;        /* whole instants only; a real host would carry the remainder */
        CHECK_CUDA(cudaMemcpyAsync(d_wide, h_wide, (size_t)n * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, NULL));
        CHECK_TDM(tdm_chan_process_ex(chan, d_wide, TDM_CHAN_IN_CS16, n, d_inst, M, TDM_CHAN_OUT_INSTANT_MAJOR, NULL));
        tdm_io io;
        memset(&io, 0, sizeof io);
        io.iq = d_inst; io.in_stride = 1; io.sample_stride = (uint32_t)M; io.count = (int32_t)(n / D); io.mem_kind = TDM_MEM_DEVICE;
        io.dibits = d_dibits; io.out_stride = S; io.out_counts = d_counts; io.out_flags = TDM_OUT_DIBITS;
        CHECK_TDM(tdm_process_io(demod, &io));                   /* state of all M chains is carried in the handle */
        CHECK_CUDA(cudaStreamSynchronize(NULL));                 /* the dibits of this chunk are in d_dibits now: hand them on here */
        wide_total += n;
    }
    fclose(f);

    tdm_metrics* m = (tdm_metrics*)malloc(sizeof(tdm_metrics) * (size_t)M);
    CHECK_TDM(tdm_get_metrics(demod, m, M));
    int locked = 0;
    for (int32_t c = 0; c < M; ++c) {
        if (m[c].sync) {
            printf("channel %d (%+.1f kHz): locked, %llu symbols, standarderr %.3f\n", c, (c <= M / 2 ? c : c - M) * 25.0,
                   (unsigned long long)m[c].n_symbols, m[c].standarderr);
            ++locked;
        }
    }
    printf("%lld wideband samples, %d channels, %d locked\n", wide_total, M, locked);
    free(m);
    cudaFreeHost(h_wide); cudaFree(d_wide); cudaFree(d_inst); cudaFree(d_dibits); cudaFree(d_counts);
    tdm_destroy(demod);
    tdm_chan_destroy(chan);
    return 0;
}
