/* Batched use of the B200 library from C: the whole of the reference's Monte-Carlo trial
 * (reference perftest/src/main.rs:9-28) for 65536 TM8192 codewords in a handful of calls.
 * Host buffers throughout (pinned, so the chunked H2D / kernel / D2H pipeline runs at PCIe rate);
 * the same entry points accept device pointers.
 *
 *   gcc examples/batch_example.c -I include -L labrador_ldpc_b200/lib -llabrador_ldpc -lm \
 *       -Wl,-rpath,$PWD/labrador_ldpc_b200/lib -o batch_example && ./batch_example
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "labrador_ldpc.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != LABRADOR_LDPC_OK) {                                                     \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, labrador_ldpc_last_error());     \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int main(void) {
    const enum labrador_ldpc_code code = LABRADOR_LDPC_CODE_TM8192;
    const size_t batch = 65536, n = labrador_ldpc_code_n(code), k = labrador_ldpc_code_k(code);
    const size_t out_len = labrador_ldpc_output_len(code);
    const double ebn0_db = 2.0, sigma2 = 1.0 / (2.0 * ((double)k / n) * pow(10.0, ebn0_db / 10.0));
    const uint64_t seed = 2026;

    CHECK(labrador_ldpc_cuda_init(NULL, 0));
    uint8_t *data = labrador_ldpc_alloc_pinned(batch * k / 8);
    uint8_t *cw = labrador_ldpc_alloc_pinned(batch * n / 8);
    int8_t *llrs = labrador_ldpc_alloc_pinned(batch * n);
    uint8_t *out = labrador_ldpc_alloc_pinned(batch * out_len);
    uint8_t *ok = labrador_ldpc_alloc_pinned(batch);
    uint32_t *iters = labrador_ldpc_alloc_pinned(batch * sizeof(uint32_t));
    uint32_t *errs = labrador_ldpc_alloc_pinned(batch * sizeof(uint32_t));
    if (!data || !cw || !llrs || !out || !ok || !iters || !errs) return 2;

    CHECK(labrador_ldpc_random_data_batch(code, seed, 0, data, batch));
    CHECK(labrador_ldpc_copy_encode_batch(code, data, cw, batch));
    /* LLR = 2 y / sigma^2, quantised to i8 as clamp(round(4 LLR), -31, 31) */
    CHECK(labrador_ldpc_awgn_batch(code, LABRADOR_LDPC_LLR_I8, cw, (float)sqrt(sigma2), (float)(8.0 / sigma2), 31, seed, 0,
                                   llrs, batch));
    CHECK(labrador_ldpc_decode_ms_i8_batch(code, llrs, out, batch, 100, ok, iters));
    CHECK(labrador_ldpc_count_errors_batch(code, out, data, errs, batch));

    unsigned long long bit_errors = 0, frame_errors = 0, it_sum = 0, ok_sum = 0;
    for (size_t f = 0; f < batch; f++) {
        bit_errors += errs[f];
        frame_errors += errs[f] != 0;
        it_sum += iters[f];
        ok_sum += ok[f];
    }
    printf("TM8192 i8 @ %.1f dB: %zu frames, %llu decoded, %llu frame errors, BER %.3e, mean iterations %.2f (%s)\n", ebn0_db,
           batch, ok_sum, frame_errors, (double)bit_errors / ((double)batch * k), (double)it_sum / batch,
           labrador_ldpc_decode_ms_kernel_name(code, LABRADOR_LDPC_LLR_I8));

    labrador_ldpc_free_pinned(data); labrador_ldpc_free_pinned(cw); labrador_ldpc_free_pinned(llrs);
    labrador_ldpc_free_pinned(out); labrador_ldpc_free_pinned(ok); labrador_ldpc_free_pinned(iters);
    labrador_ldpc_free_pinned(errs);
    labrador_ldpc_cuda_shutdown();
    return ok_sum > batch * 99 / 100 ? 0 : 3;
}
