/* Calls per second of the reference-signature single-codeword API from T host threads (the reference's usage:
 * one decoder per host thread, perftest/src/main.rs:39-45).  Development / profiles aid.
 *   gcc -O2 -pthread -Iinclude tools/calls_per_s.c -o /tmp/calls_per_s -Llabrador_ldpc_b200/lib -llabrador_ldpc \
 *       -Wl,-rpath,$PWD/labrador_ldpc_b200/lib
 *   /tmp/calls_per_s CODE THREADS CALLS_PER_THREAD */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "labrador_ldpc.h"

static enum labrador_ldpc_code code;
static int calls;
static size_t n, k, out_len;
static uint8_t *codeword;

static void *worker(void *arg) {
    int8_t *llrs = malloc(n);
    uint8_t *out = malloc(out_len);
    labrador_ldpc_hard_to_llrs_i8(code, codeword, llrs);
    for (size_t i = 0; i < n; i++) llrs[i] = (int8_t)(llrs[i] * 8);
    for (size_t i = 0; i < n; i += 7) llrs[i] = (int8_t)(-llrs[i] / 4);   /* a few weak wrong bits */
    size_t iters = 0;
    long ok = 0;
    for (int i = 0; i < calls; i++) ok += labrador_ldpc_decode_ms_i8(code, llrs, out, NULL, NULL, 50, &iters);
    *(long *)arg = ok == calls && memcmp(out, codeword, n / 8) == 0;
    free(llrs);
    free(out);
    return NULL;
}

int main(int argc, char **argv) {
    code = (enum labrador_ldpc_code)(argc > 1 ? atoi(argv[1]) : 8);
    int threads = argc > 2 ? atoi(argv[2]) : 8;
    calls = argc > 3 ? atoi(argv[3]) : 2000;
    n = labrador_ldpc_code_n(code);
    k = labrador_ldpc_code_k(code);
    out_len = labrador_ldpc_output_len(code);
    uint8_t *data = malloc(k / 8);
    codeword = malloc(n / 8);
    for (size_t i = 0; i < k / 8; i++) data[i] = (uint8_t)(i * 37 + 11);
    labrador_ldpc_copy_encode(code, data, codeword);          /* also warms the context up */
    pthread_t *th = malloc(sizeof(pthread_t) * threads);
    long *good = calloc(threads, sizeof(long));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &good[t]);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    long all = 1;
    for (int t = 0; t < threads; t++) all &= good[t];
    printf("code %d threads %d: %.0f calls/s (%.1f us per call per thread), results %s\n", (int)code, threads,
           (double)threads * calls / s, s / calls * 1e6, all ? "correct" : "WRONG");
    return all ? 0 : 1;
}
