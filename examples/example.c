/* The reference's own C example pattern (reference capi/examples/example.c: static buffers sized with the
 * LABRADOR_LDPC_* macros, encode -> corrupt -> hard_to_llrs -> decode_ms, and decode_bf), unchanged in shape,
 * linked against the B200 library.  Needs a CUDA device at run time (there is no CPU fallback).
 *
 *   gcc examples/example.c -I include -L labrador_ldpc_b200/lib -llabrador_ldpc \
 *       -Wl,-rpath,$PWD/labrador_ldpc_b200/lib -o example && ./example
 */
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "labrador_ldpc.h"

#define CODE TM2048

static uint8_t txcode[LABRADOR_LDPC_N(CODE) / 8];
static uint8_t rxcode[LABRADOR_LDPC_N(CODE) / 8];
static float llrs[LABRADOR_LDPC_N(CODE)];
static float working[LABRADOR_LDPC_MS_WORKING_LEN(CODE)];
static uint8_t working_u8[LABRADOR_LDPC_MS_WORKING_U8_LEN(CODE)];
static uint8_t bf_working[LABRADOR_LDPC_BF_WORKING_LEN(CODE)];
static uint8_t output[LABRADOR_LDPC_OUTPUT_LEN(CODE)];

int main(void) {
    const enum labrador_ldpc_code code = LABRADOR_LDPC_CODE(CODE);
    const size_t k8 = labrador_ldpc_code_k(code) / 8, n8 = labrador_ldpc_code_n(code) / 8;
    size_t iters = 0;

    for (size_t i = 0; i < k8; i++) txcode[i] = (uint8_t)(i * 37 + 11);
    labrador_ldpc_encode(code, txcode);

    memcpy(rxcode, txcode, n8);
    rxcode[3] ^= 0x28;
    rxcode[100] ^= 0x01;

    labrador_ldpc_hard_to_llrs_f32(code, rxcode, llrs);
    bool ok = labrador_ldpc_decode_ms_f32(code, llrs, output, working, working_u8, 50, &iters);
    printf("decode_ms_f32: %s after %zu iterations, data %s\n", ok ? "ok" : "FAILED", iters,
           memcmp(output, txcode, k8) == 0 ? "recovered" : "WRONG");
    if (!ok || memcmp(output, txcode, n8) != 0) return 1;

    ok = labrador_ldpc_decode_bf(code, rxcode, output, bf_working, 50, &iters);
    printf("decode_bf:     %s after %zu iterations, data %s\n", ok ? "ok" : "FAILED", iters,
           memcmp(output, txcode, k8) == 0 ? "recovered" : "WRONG");
    return ok && memcmp(output, txcode, n8) == 0 ? 0 : 1;
}
