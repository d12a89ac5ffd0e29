/* Plain-C use of the two C ABIs (include/tdm_b200.h, include/tdm_burst_b200.h): demodulate C channels from host
 * buffers, keep the decoded bits on the caller's side, feed them to the burst synchroniser, print what was found.
 *   gcc -std=c11 -I include examples/demod_then_bursts.c -L sdrpp_tetra_demodulator_b200 -ltdm_b200 -o demod_then_bursts
 * Reads interleaved float32 IQ (re, im) for ONE channel from a file; without a B200 it stops at tdm_create with
 * TDM_ERR_NO_DEVICE (there is no CPU fallback).  tests/test_capi_cpu.py builds it and checks exactly that. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "tdm_b200.h"
#include "tdm_burst_b200.h"

int main(int argc, char** argv) {
    const int32_t chunk = 32768;                       /* samples per tdm_process call, state carried in the handle */
    tdm_config cfg;
    tdm_default_config(&cfg);                          /* what src/main.cpp:35-44,78-84 passes to PI4DQPSK::init */
    tdm_handle* demod = NULL;
    int rc = tdm_create(&cfg, 1, chunk, 0, &demod);
    if (rc != TDM_OK) {
        fprintf(stderr, "tdm_create: status %d: %s\n", rc, tdm_last_error());
        return rc == TDM_ERR_NO_DEVICE ? 3 : 1;
    }
    tdm_bsync* bsync = NULL;
    const int64_t stride = tdm_max_symbols(demod, chunk);
    if (tdm_bsync_create(1, 2 * stride, 0, &bsync) != TDM_OK) { fprintf(stderr, "%s\n", tdm_last_error()); return 1; }

    FILE* f = argc > 1 ? fopen(argv[1], "rb") : NULL;
    if (!f) { fprintf(stderr, "usage: demod_then_bursts capture.f32\n"); return 2; }
    float* iq = (float*)malloc(sizeof(float) * 2 * (size_t)chunk);
    uint8_t* bits = (uint8_t*)malloc(2 * (size_t)stride);
    tdm_burst* bursts = (tdm_burst*)malloc(sizeof(tdm_burst) * (size_t)(2 * stride / 432 + 2));
    size_t got;
    long long n_bursts_total = 0;
    while ((got = fread(iq, 2 * sizeof(float), (size_t)chunk, f)) > 0) {
        int32_t n_sym = 0, n_b = 0;
        /* PI4DQPSK::process -> DQPSKSymbolExtractor::process -> BitUnpacker::process */
        if (tdm_process(demod, iq, (int64_t)got, (int32_t)got, NULL, NULL, bits, stride, &n_sym, TDM_OUT_BITS, TDM_MEM_HOST) != TDM_OK) { break; }
        /* osmotetradec::process: tetra_burst_sync_in, 432 bits per call */
        const int32_t n_bits = 2 * n_sym;
        if (tdm_bsync_in(bsync, bits, 2 * stride, NULL, n_bits, TDM_BSYNC_IN_BITS, 432, bursts, (int32_t)(n_bits / 432 + 2), &n_b, 1, TDM_MEM_HOST) != TDM_OK) { break; }
        for (int32_t i = 0; i < n_b; ++i) {
            printf("burst at bit %u: train_seq %d, tn/fn/mn %u/%u/%u\n", bursts[i].bitnum, bursts[i].train_seq, bursts[i].tn, bursts[i].fn, bursts[i].mn);
        }
        n_bursts_total += n_b;
    }
    tdm_metrics m;
    tdm_get_metrics(demod, &m, 1);
    printf("%llu samples, %llu symbols, sync %u, standarderr %.3f, %lld bursts\n", (unsigned long long)m.n_samples, (unsigned long long)m.n_symbols,
           m.sync, m.standarderr, n_bursts_total);
    fclose(f);
    free(iq); free(bits); free(bursts);
    tdm_bsync_destroy(bsync);
    tdm_destroy(demod);
    return 0;
}
