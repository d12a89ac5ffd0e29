/* Multi-GPU use of the C ABI without Python: one process per GPU, channels sharded by contiguous blocks, decoded symbol
 * streams gathered to rank 0 over NCCL (include/tdm_b200.h "multi-GPU epilogue"; BASELINE.json configs[3]).
 *
 *   gcc -std=c11 -I include -I /usr/local/cuda/include examples/sharded_gather.c \
 *       -L sdrpp_tetra_demodulator_b200 -ltdm_b200 -L /usr/local/cuda/lib64 -lcudart -o sharded_gather
 *   ./sharded_gather 0 2 /tmp/tdm.id &  ./sharded_gather 1 2 /tmp/tdm.id        (rank, world, rendezvous file)
 *
 * Rank r demodulates channels [r C, (r+1) C) of a synthetic capture generated on its own GPU, the slicer packs four
 * dibits per byte, tdm_gather_packed moves packed rows + counts to rank 0, which unpacks them into the dibit stream
 * DQPSKSymbolExtractor would have produced (src/dsp/dqpsk_sym_extr.cpp:4-55) and compares every channel with what was
 * transmitted.  The 128-byte NCCL id travels through a file here; any channel the host application has will do. */
#define _DEFAULT_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <cuda_runtime_api.h>

#include "tdm_b200.h"

#define CHECK_TDM(call) do { int rc_ = (call); if (rc_ != TDM_OK) { fprintf(stderr, "%s: status %d: %s\n", #call, rc_, tdm_last_error()); return rc_ == TDM_ERR_NO_DEVICE ? 3 : 1; } } while (0)
#define CHECK_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: sharded_gather RANK WORLD ID_FILE [channels_per_rank] [samples]\n"); return 2; }
    const int32_t rank = atoi(argv[1]), world = atoi(argv[2]);
    const char* id_file = argv[3];
    const int32_t C = argc > 4 ? atoi(argv[4]) : 64;
    const int32_t N = argc > 5 ? atoi(argv[5]) : 60000;

    tdm_config cfg;
    tdm_default_config(&cfg);
    tdm_handle* demod = NULL;
    CHECK_TDM(tdm_create(&cfg, C, N, rank, &demod));                 /* device = rank; no B200: TDM_ERR_NO_DEVICE, no fallback */
    CHECK_CUDA(cudaSetDevice(rank));

    /* rendezvous: rank 0 makes the id, the others wait for the file */
    uint8_t id[TDM_COMM_ID_BYTES];
    if (rank == 0) {
        CHECK_TDM(tdm_comm_unique_id(id));
        char tmp[1024];
        snprintf(tmp, sizeof tmp, "%s.tmp", id_file);
        FILE* f = fopen(tmp, "wb");
        if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "cannot write %s\n", tmp); return 1; }
        fclose(f);
        rename(tmp, id_file);
    } else {
        FILE* f = NULL;
        for (int tries = 0; tries < 600 && !(f = fopen(id_file, "rb")); ++tries) { usleep(100000); }
        if (!f || fread(id, 1, sizeof id, f) != sizeof id) { fprintf(stderr, "no id in %s\n", id_file); return 1; }
        fclose(f);
    }
    tdm_comm* comm = NULL;
    CHECK_TDM(tdm_comm_create(id, rank, world, rank, &comm));

    /* this rank's shard of the capture, generated where it is consumed */
    const int64_t S = tdm_max_symbols(demod, N), PS = (S + 3) / 4, TXS = N / 2 + 64;
    float* iq; uint8_t *tx, *packed, *packed_all = NULL, *dibits = NULL; int32_t *counts, *counts_all = NULL;
    CHECK_CUDA(cudaMalloc((void**)&iq, sizeof(float) * 2 * (size_t)C * (size_t)N));
    CHECK_CUDA(cudaMalloc((void**)&tx, (size_t)C * (size_t)TXS));
    CHECK_CUDA(cudaMalloc((void**)&packed, (size_t)C * (size_t)PS));
    CHECK_CUDA(cudaMalloc((void**)&counts, sizeof(int32_t) * (size_t)C));
    if (rank == 0) {
        CHECK_CUDA(cudaMalloc((void**)&packed_all, (size_t)world * (size_t)C * (size_t)PS));
        CHECK_CUDA(cudaMalloc((void**)&counts_all, sizeof(int32_t) * (size_t)world * (size_t)C));
        CHECK_CUDA(cudaMalloc((void**)&dibits, (size_t)world * (size_t)C * (size_t)S));
    }
    tdm_synth_params sp;
    memset(&sp, 0, sizeof sp);
    sp.snr_db = 30.0; sp.max_freq_off_hz = 300.0; sp.min_amp = 0.05; sp.max_amp = 2.0; sp.seed_data = 12345; sp.seed_noise = 777;
    CHECK_TDM(tdm_synth_capture(rank, NULL, &sp, C, N, N, rank * C, iq, tx, TXS));

    /* PI4DQPSK::process .. BitUnpacker::process for the whole shard, packed output only */
    tdm_io io;
    memset(&io, 0, sizeof io);
    io.iq = iq; io.in_stride = N; io.count = N; io.mem_kind = TDM_MEM_DEVICE;
    io.packed = packed; io.packed_stride = PS; io.out_stride = S; io.out_counts = counts; io.out_flags = TDM_OUT_PACKED;
    CHECK_TDM(tdm_process_io(demod, &io));
    CHECK_CUDA(cudaDeviceSynchronize());                              /* the handle's stream; the gather below uses the default stream */
    CHECK_TDM(tdm_gather_packed(comm, 0, C, packed, PS, counts, packed_all, counts_all, NULL));
    CHECK_CUDA(cudaDeviceSynchronize());

    int bad_channels = 0;
    if (rank == 0) {
        CHECK_TDM(tdm_unpack_dibits(demod, packed_all, PS, counts_all, world * C, dibits, S, NULL, 0, S));
        CHECK_CUDA(cudaDeviceSynchronize());
        /* rank 0 can check its own rows against what it transmitted: the receiver's dibits are the transmitted ones
         * delayed by the filters (a constant lag), once the loops have locked -- the second half is enough here */
        uint8_t* h_d = (uint8_t*)malloc((size_t)C * (size_t)S);
        uint8_t* h_t = (uint8_t*)malloc((size_t)C * (size_t)TXS);
        int32_t* h_c = (int32_t*)malloc(sizeof(int32_t) * (size_t)world * (size_t)C);
        CHECK_CUDA(cudaMemcpy(h_d, dibits, (size_t)C * (size_t)S, cudaMemcpyDeviceToHost));
        CHECK_CUDA(cudaMemcpy(h_t, tx, (size_t)C * (size_t)TXS, cudaMemcpyDeviceToHost));
        CHECK_CUDA(cudaMemcpy(h_c, counts_all, sizeof(int32_t) * (size_t)world * (size_t)C, cudaMemcpyDeviceToHost));
        long long total = 0;
        for (int32_t r = 0; r < world * C; ++r) { total += h_c[r]; }
        for (int32_t c = 0; c < C; ++c) {
            const int32_t n = h_c[c];
            int best = -1;
            for (int lag = 8; lag < 32 && best != 0; ++lag) {
                int errs = 0;
                for (int32_t k = n / 2; k < n; ++k) { errs += h_d[(size_t)c * (size_t)S + (size_t)k] != h_t[(size_t)c * (size_t)TXS + (size_t)(k - lag)]; }
                if (best < 0 || errs < best) { best = errs; }
            }
            bad_channels += best != 0;
        }
        printf("rank 0 holds %d channels, %lld symbols; %d of its own %d channels differ from the transmitted dibits after lock\n",
               world * C, total, bad_channels, C);
        free(h_d); free(h_t); free(h_c);
    }
    tdm_comm_destroy(comm);
    tdm_destroy(demod);
    cudaFree(iq); cudaFree(tx); cudaFree(packed); cudaFree(counts); cudaFree(packed_all); cudaFree(counts_all); cudaFree(dibits);
    return bad_channels ? 4 : 0;
}
