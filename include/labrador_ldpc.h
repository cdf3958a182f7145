/* labrador_ldpc.h -- C ABI of the B200-native batched LDPC codec.
 *
 * Drop-in for the C API of adamgreig/labrador-ldpc v1.2.1
 * (reference: capi/include/labrador_ldpc.h:19-244, capi/src/lib.rs:15-179):
 * the same enum, the same 21 entry points with the same signatures and
 * buffer contracts, plus `_batch` / `_batch_async` entry points that decode,
 * encode or convert many independent codewords per call on the GPU.
 *
 * Every compute entry point runs hand-written sm_100a CUDA kernels; there is
 * no CPU fallback.  If no CUDA device is usable the reference-signature
 * functions abort() with a message on stderr (they have no error channel; the
 * reference's own failure mode is a panic handler that spins forever,
 * capi/src/lib.rs:9-13) and the `_batch` functions return a negative
 * LABRADOR_LDPC_ERR_* code.
 *
 * Only plain pointers and sizes cross this boundary (no C++/torch types).
 */
#ifndef LABRADOR_LDPC_B200_CAPI
#define LABRADOR_LDPC_B200_CAPI

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Replaces: enum labrador_ldpc_code, capi/include/labrador_ldpc.h:19-29
 * (= #[repr(C)] enum LDPCCode, src/codes/mod.rs:37-66). */
enum labrador_ldpc_code {
    LABRADOR_LDPC_CODE_TC128 = 0,  /* n=128  k=64   r=1/2 */
    LABRADOR_LDPC_CODE_TC256 = 1,  /* n=256  k=128  r=1/2 */
    LABRADOR_LDPC_CODE_TC512 = 2,  /* n=512  k=256  r=1/2 */
    LABRADOR_LDPC_CODE_TM1280 = 3, /* n=1280 k=1024 r=4/5 */
    LABRADOR_LDPC_CODE_TM1536 = 4, /* n=1536 k=1024 r=2/3 */
    LABRADOR_LDPC_CODE_TM2048 = 5, /* n=2048 k=1024 r=1/2 */
    LABRADOR_LDPC_CODE_TM5120 = 6, /* n=5120 k=4096 r=4/5 */
    LABRADOR_LDPC_CODE_TM6144 = 7, /* n=6144 k=4096 r=2/3 */
    LABRADOR_LDPC_CODE_TM8192 = 8, /* n=8192 k=4096 r=1/2 */
    /* Extension (not in the reference's enum): the k = 16384 codes of CCSDS 131.0-B.  The reference carries their
     * parity-check constants (src/codes/compact_parity_checks.rs:84-96, PHI_J_K_M4096 / M8192, selected at
     * src/codes/mod.rs:473-476) but leaves the codes out for want of generator matrices (src/lib.rs:81-83).  Here they
     * are decoded like every TM code and encoded through the sparse parity-check matrix, which needs no generator. */
    LABRADOR_LDPC_CODE_TM20480 = 9,  /* n=20480 k=16384 r=4/5 */
    LABRADOR_LDPC_CODE_TM24576 = 10, /* n=24576 k=16384 r=2/3 */
    LABRADOR_LDPC_CODE_TM32768 = 11, /* n=32768 k=16384 r=1/2 */
};
#define LABRADOR_LDPC_NUM_CODES 12
#define LABRADOR_LDPC_NUM_REFERENCE_CODES 9

/* ---------------------------------------------------------------------------
 * Compile-time sizes.  Replaces capi/include/labrador_ldpc.h:42-115.
 * Usable as LABRADOR_LDPC_N_TC512 or LABRADOR_LDPC_N(TC512) / N(CODE).
 * Everything is derived from (n, k, punctured bits p, edge count E):
 *   bf_working = n+p, ms_working = 2E+3n+3p-2k, ms_working_u8 = (n+p-k)/8,
 *   output = (n+p)/8                       (src/decoder.rs:93-116)
 *
 * Deliberate differences from the reference header:
 *   - the reference defines LABRADOR_LDPC_N_TM6144 as (6140), which
 *     under-allocates every buffer sized with it; here it is 6144;
 *   - the reference spells four TM6144 macros "..._TM6140"; both spellings
 *     are defined here so existing callers keep compiling.
 * ------------------------------------------------------------------------- */
#define LABRADOR_LDPC_CODE_(CODE) LABRADOR_LDPC_CODE_##CODE
#define LABRADOR_LDPC_CODE(CODE) LABRADOR_LDPC_CODE_(CODE)

#define LABRADOR_LDPC_N_TC128 (128)
#define LABRADOR_LDPC_K_TC128 (64)
#define LABRADOR_LDPC_P_TC128 (0)
#define LABRADOR_LDPC_E_TC128 (512)
#define LABRADOR_LDPC_N_TC256 (256)
#define LABRADOR_LDPC_K_TC256 (128)
#define LABRADOR_LDPC_P_TC256 (0)
#define LABRADOR_LDPC_E_TC256 (1024)
#define LABRADOR_LDPC_N_TC512 (512)
#define LABRADOR_LDPC_K_TC512 (256)
#define LABRADOR_LDPC_P_TC512 (0)
#define LABRADOR_LDPC_E_TC512 (2048)
#define LABRADOR_LDPC_N_TM1280 (1280)
#define LABRADOR_LDPC_K_TM1280 (1024)
#define LABRADOR_LDPC_P_TM1280 (128)
#define LABRADOR_LDPC_E_TM1280 (4992)
#define LABRADOR_LDPC_N_TM1536 (1536)
#define LABRADOR_LDPC_K_TM1536 (1024)
#define LABRADOR_LDPC_P_TM1536 (256)
#define LABRADOR_LDPC_E_TM1536 (5888)
#define LABRADOR_LDPC_N_TM2048 (2048)
#define LABRADOR_LDPC_K_TM2048 (1024)
#define LABRADOR_LDPC_P_TM2048 (512)
#define LABRADOR_LDPC_E_TM2048 (7680)
#define LABRADOR_LDPC_N_TM5120 (5120)
#define LABRADOR_LDPC_K_TM5120 (4096)
#define LABRADOR_LDPC_P_TM5120 (512)
#define LABRADOR_LDPC_E_TM5120 (19968)
#define LABRADOR_LDPC_N_TM6144 (6144)
#define LABRADOR_LDPC_K_TM6144 (4096)
#define LABRADOR_LDPC_P_TM6144 (1024)
#define LABRADOR_LDPC_E_TM6144 (23552)
#define LABRADOR_LDPC_N_TM6140 LABRADOR_LDPC_N_TM6144
#define LABRADOR_LDPC_K_TM6140 LABRADOR_LDPC_K_TM6144
#define LABRADOR_LDPC_P_TM6140 LABRADOR_LDPC_P_TM6144
#define LABRADOR_LDPC_E_TM6140 LABRADOR_LDPC_E_TM6144
#define LABRADOR_LDPC_N_TM8192 (8192)
#define LABRADOR_LDPC_K_TM8192 (4096)
#define LABRADOR_LDPC_P_TM8192 (2048)
#define LABRADOR_LDPC_E_TM8192 (30720)
#define LABRADOR_LDPC_N_TM20480 (20480)
#define LABRADOR_LDPC_K_TM20480 (16384)
#define LABRADOR_LDPC_P_TM20480 (2048)
#define LABRADOR_LDPC_E_TM20480 (79872)
#define LABRADOR_LDPC_N_TM24576 (24576)
#define LABRADOR_LDPC_K_TM24576 (16384)
#define LABRADOR_LDPC_P_TM24576 (4096)
#define LABRADOR_LDPC_E_TM24576 (94208)
#define LABRADOR_LDPC_N_TM32768 (32768)
#define LABRADOR_LDPC_K_TM32768 (16384)
#define LABRADOR_LDPC_P_TM32768 (8192)
#define LABRADOR_LDPC_E_TM32768 (122880)

#define LABRADOR_LDPC_N_(CODE) LABRADOR_LDPC_N_##CODE
#define LABRADOR_LDPC_N(CODE) LABRADOR_LDPC_N_(CODE)
#define LABRADOR_LDPC_K_(CODE) LABRADOR_LDPC_K_##CODE
#define LABRADOR_LDPC_K(CODE) LABRADOR_LDPC_K_(CODE)
#define LABRADOR_LDPC_P_(CODE) LABRADOR_LDPC_P_##CODE
#define LABRADOR_LDPC_P(CODE) LABRADOR_LDPC_P_(CODE)
#define LABRADOR_LDPC_E_(CODE) LABRADOR_LDPC_E_##CODE
#define LABRADOR_LDPC_E(CODE) LABRADOR_LDPC_E_(CODE)

#define LABRADOR_LDPC_BF_WORKING_LEN(CODE) (LABRADOR_LDPC_N(CODE) + LABRADOR_LDPC_P(CODE))
#define LABRADOR_LDPC_MS_WORKING_LEN(CODE)                                                  \
    (2 * LABRADOR_LDPC_E(CODE) + 3 * LABRADOR_LDPC_N(CODE) + 3 * LABRADOR_LDPC_P(CODE) -    \
     2 * LABRADOR_LDPC_K(CODE))
#define LABRADOR_LDPC_MS_WORKING_U8_LEN(CODE) \
    ((LABRADOR_LDPC_N(CODE) + LABRADOR_LDPC_P(CODE) - LABRADOR_LDPC_K(CODE)) / 8)
#define LABRADOR_LDPC_OUTPUT_LEN(CODE) ((LABRADOR_LDPC_N(CODE) + LABRADOR_LDPC_P(CODE)) / 8)

/* Per-code spellings (LABRADOR_LDPC_MS_WORKING_LEN_TC128, ...) as in the reference. */
#define LABRADOR_LDPC_BF_WORKING_LEN_TC128 LABRADOR_LDPC_BF_WORKING_LEN(TC128)
#define LABRADOR_LDPC_BF_WORKING_LEN_TC256 LABRADOR_LDPC_BF_WORKING_LEN(TC256)
#define LABRADOR_LDPC_BF_WORKING_LEN_TC512 LABRADOR_LDPC_BF_WORKING_LEN(TC512)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM1280 LABRADOR_LDPC_BF_WORKING_LEN(TM1280)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM1536 LABRADOR_LDPC_BF_WORKING_LEN(TM1536)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM2048 LABRADOR_LDPC_BF_WORKING_LEN(TM2048)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM5120 LABRADOR_LDPC_BF_WORKING_LEN(TM5120)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM6144 LABRADOR_LDPC_BF_WORKING_LEN(TM6144)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM6140 LABRADOR_LDPC_BF_WORKING_LEN(TM6144)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM8192 LABRADOR_LDPC_BF_WORKING_LEN(TM8192)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM20480 LABRADOR_LDPC_BF_WORKING_LEN(TM20480)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM24576 LABRADOR_LDPC_BF_WORKING_LEN(TM24576)
#define LABRADOR_LDPC_BF_WORKING_LEN_TM32768 LABRADOR_LDPC_BF_WORKING_LEN(TM32768)
#define LABRADOR_LDPC_MS_WORKING_LEN_TC128 LABRADOR_LDPC_MS_WORKING_LEN(TC128)
#define LABRADOR_LDPC_MS_WORKING_LEN_TC256 LABRADOR_LDPC_MS_WORKING_LEN(TC256)
#define LABRADOR_LDPC_MS_WORKING_LEN_TC512 LABRADOR_LDPC_MS_WORKING_LEN(TC512)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM1280 LABRADOR_LDPC_MS_WORKING_LEN(TM1280)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM1536 LABRADOR_LDPC_MS_WORKING_LEN(TM1536)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM2048 LABRADOR_LDPC_MS_WORKING_LEN(TM2048)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM5120 LABRADOR_LDPC_MS_WORKING_LEN(TM5120)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM6144 LABRADOR_LDPC_MS_WORKING_LEN(TM6144)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM6140 LABRADOR_LDPC_MS_WORKING_LEN(TM6144)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM8192 LABRADOR_LDPC_MS_WORKING_LEN(TM8192)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM20480 LABRADOR_LDPC_MS_WORKING_LEN(TM20480)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM24576 LABRADOR_LDPC_MS_WORKING_LEN(TM24576)
#define LABRADOR_LDPC_MS_WORKING_LEN_TM32768 LABRADOR_LDPC_MS_WORKING_LEN(TM32768)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TC128 LABRADOR_LDPC_MS_WORKING_U8_LEN(TC128)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TC256 LABRADOR_LDPC_MS_WORKING_U8_LEN(TC256)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TC512 LABRADOR_LDPC_MS_WORKING_U8_LEN(TC512)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM1280 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM1280)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM1536 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM1536)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM2048 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM2048)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM5120 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM5120)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM6144 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM6144)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM6140 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM6144)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM8192 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM8192)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM20480 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM20480)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM24576 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM24576)
#define LABRADOR_LDPC_MS_WORKING_U8_LEN_TM32768 LABRADOR_LDPC_MS_WORKING_U8_LEN(TM32768)
#define LABRADOR_LDPC_OUTPUT_LEN_TC128 LABRADOR_LDPC_OUTPUT_LEN(TC128)
#define LABRADOR_LDPC_OUTPUT_LEN_TC256 LABRADOR_LDPC_OUTPUT_LEN(TC256)
#define LABRADOR_LDPC_OUTPUT_LEN_TC512 LABRADOR_LDPC_OUTPUT_LEN(TC512)
#define LABRADOR_LDPC_OUTPUT_LEN_TM1280 LABRADOR_LDPC_OUTPUT_LEN(TM1280)
#define LABRADOR_LDPC_OUTPUT_LEN_TM1536 LABRADOR_LDPC_OUTPUT_LEN(TM1536)
#define LABRADOR_LDPC_OUTPUT_LEN_TM2048 LABRADOR_LDPC_OUTPUT_LEN(TM2048)
#define LABRADOR_LDPC_OUTPUT_LEN_TM5120 LABRADOR_LDPC_OUTPUT_LEN(TM5120)
#define LABRADOR_LDPC_OUTPUT_LEN_TM6144 LABRADOR_LDPC_OUTPUT_LEN(TM6144)
#define LABRADOR_LDPC_OUTPUT_LEN_TM6140 LABRADOR_LDPC_OUTPUT_LEN(TM6144)
#define LABRADOR_LDPC_OUTPUT_LEN_TM8192 LABRADOR_LDPC_OUTPUT_LEN(TM8192)
#define LABRADOR_LDPC_OUTPUT_LEN_TM20480 LABRADOR_LDPC_OUTPUT_LEN(TM20480)
#define LABRADOR_LDPC_OUTPUT_LEN_TM24576 LABRADOR_LDPC_OUTPUT_LEN(TM24576)
#define LABRADOR_LDPC_OUTPUT_LEN_TM32768 LABRADOR_LDPC_OUTPUT_LEN(TM32768)

/* ===========================================================================
 * Part 1 -- the reference's 21 entry points (single codeword per call).
 * Same names, argument meaning and buffer ownership as the reference: the
 * caller owns every buffer; `working*` arguments are accepted for source
 * compatibility and are not touched (scratch lives in GPU shared memory).
 * Each call runs a batch of one on the current CUDA context (lazy init).
 * ========================================================================= */

/* Replaces labrador_ldpc_code_n, capi/src/lib.rs:15-18 (LDPCCode::n, src/codes/mod.rs:382). */
size_t labrador_ldpc_code_n(enum labrador_ldpc_code code);
/* Replaces labrador_ldpc_code_k, capi/src/lib.rs:20-23 (src/codes/mod.rs:387). */
size_t labrador_ldpc_code_k(enum labrador_ldpc_code code);
/* Replaces labrador_ldpc_bf_working_len, capi/src/lib.rs:48-51 (src/decoder.rs:93). */
size_t labrador_ldpc_bf_working_len(enum labrador_ldpc_code code);
/* Replaces labrador_ldpc_ms_working_u8_len, capi/src/lib.rs:58-61 (src/decoder.rs:107). */
size_t labrador_ldpc_ms_working_u8_len(enum labrador_ldpc_code code);
/* Replaces labrador_ldpc_ms_working_len, capi/src/lib.rs:53-56 (src/decoder.rs:100). */
size_t labrador_ldpc_ms_working_len(enum labrador_ldpc_code code);
/* Replaces labrador_ldpc_output_len, capi/src/lib.rs:63-66 (src/decoder.rs:114). */
size_t labrador_ldpc_output_len(enum labrador_ldpc_code code);

/* Replaces labrador_ldpc_encode, capi/src/lib.rs:25-34 (LDPCCode::encode, src/encoder.rs:293).
 * `codeword` is n/8 bytes; first k/8 bytes in, last (n-k)/8 bytes out. */
void labrador_ldpc_encode(enum labrador_ldpc_code code, uint8_t *codeword);
/* Replaces labrador_ldpc_copy_encode, capi/src/lib.rs:36-46 (src/encoder.rs:309). */
void labrador_ldpc_copy_encode(enum labrador_ldpc_code code, const uint8_t *data, uint8_t *codeword);

/* Replaces labrador_ldpc_decode_bf, capi/src/lib.rs:68-81 (LDPCCode::decode_bf, src/decoder.rs:243).
 * input n/8 bytes, output (n+p)/8 bytes; iters_run may be NULL. */
bool labrador_ldpc_decode_bf(enum labrador_ldpc_code code, const uint8_t *input, uint8_t *output,
                             uint8_t *working, size_t max_iters, size_t *iters_run);

/* Replace labrador_ldpc_decode_ms_{i8,i16,f32,f64}, capi/src/lib.rs:97-127
 * (LDPCCode::decode_ms<T>, src/decoder.rs:347).  llrs: n elements, positive = bit 0.
 * f32 / f64: LLRs must not be NaN.  The reference orders magnitudes with `<` (src/decoder.rs:430-435), which skips
 * a NaN; the kernels use the hardware minimum, which treats a NaN operand differently, so results for NaN input
 * are unspecified (every other value, +-inf and +-0 included, follows the reference bit for bit).  The soft front
 * ends (labrador_ldpc_decode_ms_*_soft_batch) map NaN to 0 before decoding. */
bool labrador_ldpc_decode_ms_i8(enum labrador_ldpc_code code, const int8_t *llrs, uint8_t *output,
                                int8_t *working, uint8_t *working_u8, size_t max_iters,
                                size_t *iters_run);
bool labrador_ldpc_decode_ms_i16(enum labrador_ldpc_code code, const int16_t *llrs, uint8_t *output,
                                 int16_t *working, uint8_t *working_u8, size_t max_iters,
                                 size_t *iters_run);
bool labrador_ldpc_decode_ms_f32(enum labrador_ldpc_code code, const float *llrs, uint8_t *output,
                                 float *working, uint8_t *working_u8, size_t max_iters,
                                 size_t *iters_run);
bool labrador_ldpc_decode_ms_f64(enum labrador_ldpc_code code, const double *llrs, uint8_t *output,
                                 double *working, uint8_t *working_u8, size_t max_iters,
                                 size_t *iters_run);
/* Extension: the Rust API also implements DecodeFrom for i32 (src/decoder.rs:60-68). */
bool labrador_ldpc_decode_ms_i32(enum labrador_ldpc_code code, const int32_t *llrs, uint8_t *output,
                                 int32_t *working, uint8_t *working_u8, size_t max_iters,
                                 size_t *iters_run);

/* Replace labrador_ldpc_hard_to_llrs_*, capi/src/lib.rs:129-153 (src/decoder.rs:484). */
void labrador_ldpc_hard_to_llrs_i8(enum labrador_ldpc_code code, const uint8_t *input, int8_t *llrs);
void labrador_ldpc_hard_to_llrs_i16(enum labrador_ldpc_code code, const uint8_t *input, int16_t *llrs);
void labrador_ldpc_hard_to_llrs_f32(enum labrador_ldpc_code code, const uint8_t *input, float *llrs);
void labrador_ldpc_hard_to_llrs_f64(enum labrador_ldpc_code code, const uint8_t *input, double *llrs);
void labrador_ldpc_hard_to_llrs_i32(enum labrador_ldpc_code code, const uint8_t *input, int32_t *llrs);

/* Replace labrador_ldpc_llrs_to_hard_*, capi/src/lib.rs:155-179 (src/decoder.rs:498). */
void labrador_ldpc_llrs_to_hard_i8(enum labrador_ldpc_code code, const int8_t *llrs, uint8_t *output);
void labrador_ldpc_llrs_to_hard_i16(enum labrador_ldpc_code code, const int16_t *llrs, uint8_t *output);
void labrador_ldpc_llrs_to_hard_f32(enum labrador_ldpc_code code, const float *llrs, uint8_t *output);
void labrador_ldpc_llrs_to_hard_f64(enum labrador_ldpc_code code, const double *llrs, uint8_t *output);
void labrador_ldpc_llrs_to_hard_i32(enum labrador_ldpc_code code, const int32_t *llrs, uint8_t *output);

/* ===========================================================================
 * Part 2 -- batched entry points (new).
 *
 * Layouts are frame-major and contiguous: frame f of a [batch][len] buffer
 * starts at element f*len.  Pointers may be host memory (pageable or pinned)
 * or device memory; all pointer arguments of one call must be of the same
 * kind.  Host buffers are streamed through the GPU in chunks with copies
 * overlapped with the kernels; device buffers are used in place.  If the
 * library was initialised with several devices, host-pointer batches are
 * split into contiguous per-device shards (independent codewords, no
 * collective).  The device owns all scratch.
 *
 * Return value: 0 on success, negative LABRADOR_LDPC_ERR_* otherwise
 * (labrador_ldpc_last_error() gives the text).  batch == 0 is a no-op.
 * `success` receives 1/0 per frame, `iters_run` (nullable) the reference's
 * returned iteration count per frame (src/decoder.rs:462,474,289,300).
 * ========================================================================= */
#define LABRADOR_LDPC_OK 0
#define LABRADOR_LDPC_ERR_BAD_CODE (-1)
#define LABRADOR_LDPC_ERR_NULL_POINTER (-2)
#define LABRADOR_LDPC_ERR_CUDA (-3)
#define LABRADOR_LDPC_ERR_MIXED_POINTERS (-4)
#define LABRADOR_LDPC_ERR_BAD_ARGUMENT (-5)

/* Select the devices to use (NULL / 0 = current device only).  Builds the
 * device-resident edge, circulant and generator tables (the one-time
 * expansion of src/codes/mod.rs:275-362,444-494).  Idempotent; called lazily
 * by every compute entry point. */
int labrador_ldpc_cuda_init(const int *devices, int n_devices);
void labrador_ldpc_cuda_shutdown(void);
int labrador_ldpc_cuda_device_count(void);
const char *labrador_ldpc_last_error(void);
const char *labrador_ldpc_version(void);

/* Pinned host buffers for callers that want full-speed PCIe streaming. */
void *labrador_ldpc_alloc_pinned(size_t bytes);
void labrador_ldpc_free_pinned(void *ptr);

/* Batched LDPCCode::decode_ms<T> (src/decoder.rs:347-475). */
int labrador_ldpc_decode_ms_i8_batch(enum labrador_ldpc_code code, const int8_t *llrs,
                                     uint8_t *output, size_t batch, size_t max_iters,
                                     uint8_t *success, uint32_t *iters_run);
int labrador_ldpc_decode_ms_i16_batch(enum labrador_ldpc_code code, const int16_t *llrs,
                                      uint8_t *output, size_t batch, size_t max_iters,
                                      uint8_t *success, uint32_t *iters_run);
int labrador_ldpc_decode_ms_i32_batch(enum labrador_ldpc_code code, const int32_t *llrs,
                                      uint8_t *output, size_t batch, size_t max_iters,
                                      uint8_t *success, uint32_t *iters_run);
int labrador_ldpc_decode_ms_f32_batch(enum labrador_ldpc_code code, const float *llrs,
                                      uint8_t *output, size_t batch, size_t max_iters,
                                      uint8_t *success, uint32_t *iters_run);
int labrador_ldpc_decode_ms_f64_batch(enum labrador_ldpc_code code, const double *llrs,
                                      uint8_t *output, size_t batch, size_t max_iters,
                                      uint8_t *success, uint32_t *iters_run);

/* Batched LDPCCode::decode_bf incl. the erasure pre-pass (src/decoder.rs:144-301). */
int labrador_ldpc_decode_bf_batch(enum labrador_ldpc_code code, const uint8_t *input,
                                  uint8_t *output, size_t batch, size_t max_iters,
                                  uint8_t *success, uint32_t *iters_run);

/* Batched LDPCCode::encode / copy_encode (src/encoder.rs:292-315). */
int labrador_ldpc_encode_batch(enum labrador_ldpc_code code, uint8_t *codewords, size_t batch);
int labrador_ldpc_copy_encode_batch(enum labrador_ldpc_code code, const uint8_t *data,
                                    uint8_t *codewords, size_t batch);

/* Batched converters (src/decoder.rs:484-509). */
int labrador_ldpc_hard_to_llrs_i8_batch(enum labrador_ldpc_code code, const uint8_t *input, int8_t *llrs, size_t batch);
int labrador_ldpc_hard_to_llrs_i16_batch(enum labrador_ldpc_code code, const uint8_t *input, int16_t *llrs, size_t batch);
int labrador_ldpc_hard_to_llrs_i32_batch(enum labrador_ldpc_code code, const uint8_t *input, int32_t *llrs, size_t batch);
int labrador_ldpc_hard_to_llrs_f32_batch(enum labrador_ldpc_code code, const uint8_t *input, float *llrs, size_t batch);
int labrador_ldpc_hard_to_llrs_f64_batch(enum labrador_ldpc_code code, const uint8_t *input, double *llrs, size_t batch);
int labrador_ldpc_llrs_to_hard_i8_batch(enum labrador_ldpc_code code, const int8_t *llrs, uint8_t *output, size_t batch);
int labrador_ldpc_llrs_to_hard_i16_batch(enum labrador_ldpc_code code, const int16_t *llrs, uint8_t *output, size_t batch);
int labrador_ldpc_llrs_to_hard_i32_batch(enum labrador_ldpc_code code, const int32_t *llrs, uint8_t *output, size_t batch);
int labrador_ldpc_llrs_to_hard_f32_batch(enum labrador_ldpc_code code, const float *llrs, uint8_t *output, size_t batch);
int labrador_ldpc_llrs_to_hard_f64_batch(enum labrador_ldpc_code code, const double *llrs, uint8_t *output, size_t batch);

/* ---------------------------------------------------------------------------
 * Stream-ordered variants: all pointers are DEVICE pointers on the current
 * device; the work is enqueued on `cuda_stream` (a cudaStream_t passed as
 * void*, NULL = default stream) and the call returns without synchronising.
 * `llr_type` is one of LABRADOR_LDPC_LLR_*.
 * ------------------------------------------------------------------------- */
#define LABRADOR_LDPC_LLR_I8 0
#define LABRADOR_LDPC_LLR_I16 1
#define LABRADOR_LDPC_LLR_I32 2
#define LABRADOR_LDPC_LLR_F32 3
#define LABRADOR_LDPC_LLR_F64 4

int labrador_ldpc_decode_ms_batch_async(enum labrador_ldpc_code code, int llr_type, const void *llrs,
                                        uint8_t *output, size_t batch, size_t max_iters,
                                        uint8_t *success, uint32_t *iters_run, void *cuda_stream);
int labrador_ldpc_decode_bf_batch_async(enum labrador_ldpc_code code, const uint8_t *input,
                                        uint8_t *output, size_t batch, size_t max_iters,
                                        uint8_t *success, uint32_t *iters_run, void *cuda_stream);
int labrador_ldpc_copy_encode_batch_async(enum labrador_ldpc_code code, const uint8_t *data,
                                          uint8_t *codewords, size_t batch, void *cuda_stream);
int labrador_ldpc_hard_to_llrs_batch_async(enum labrador_ldpc_code code, int llr_type,
                                           const uint8_t *input, void *llrs, size_t batch,
                                           void *cuda_stream);
int labrador_ldpc_llrs_to_hard_batch_async(enum labrador_ldpc_code code, int llr_type,
                                           const void *llrs, uint8_t *output, size_t batch,
                                           void *cuda_stream);

/* ---------------------------------------------------------------------------
 * Fused front ends of decode_ms (not in the reference).
 *
 * The reference's callers convert before they decode: hard_to_llrs
 * (src/decoder.rs:484-493; capi/examples/example.c, src/lib.rs:26-49) or a
 * float -> i8/i16 quantiser with headroom (src/decoder.rs:337-340; soft values
 * as built in perftest/src/main.rs:13-18).  These entry points do that
 * conversion while the decoder loads its channel LLRs, so the input crosses
 * HBM once.  Results are bit-identical to the two-step sequence
 *   quantise / hard_to_llrs  ->  decode_ms_{i8,i16}.
 *
 *   _soft_:  soft[B][n] floats;  llr = clamp(rint(soft * scale), -limit, +limit)
 *            with round-half-to-even, a NaN product counted as 0 (an erasure);
 *            limit in 1..127 (i8) / 1..32767 (i16), scale finite.
 *   _hard_:  input[B][n/8] bit-packed hard decisions, MSB first;
 *            bit 1 -> -1, bit 0 -> +1 (exactly hard_to_llrs::<i8>).
 * Outputs, `success` and `iters_run` as for the _batch decoders above.
 * labrador_ldpc_quantise_* is the stand-alone quantiser (same arithmetic).
 * ------------------------------------------------------------------------- */
#define LABRADOR_LDPC_FRONT_NONE 0
#define LABRADOR_LDPC_FRONT_SOFT_F32 1
#define LABRADOR_LDPC_FRONT_HARD 2

int labrador_ldpc_decode_ms_i8_soft_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                          uint8_t *output, size_t batch, size_t max_iters,
                                          uint8_t *success, uint32_t *iters_run);
int labrador_ldpc_decode_ms_i16_soft_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                           uint8_t *output, size_t batch, size_t max_iters,
                                           uint8_t *success, uint32_t *iters_run);
int labrador_ldpc_decode_ms_i8_hard_batch(enum labrador_ldpc_code code, const uint8_t *input, uint8_t *output,
                                          size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters_run);
/* Stream-ordered form: `front` is LABRADOR_LDPC_FRONT_*, `input` holds what that front end reads. */
int labrador_ldpc_decode_ms_front_batch_async(enum labrador_ldpc_code code, int llr_type, int front,
                                              const void *input, float scale, int limit, uint8_t *output,
                                              size_t batch, size_t max_iters, uint8_t *success,
                                              uint32_t *iters_run, void *cuda_stream);
int labrador_ldpc_quantise_i8_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                    int8_t *llrs, size_t batch);
int labrador_ldpc_quantise_i16_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                     int16_t *llrs, size_t batch);
int labrador_ldpc_quantise_batch_async(enum labrador_ldpc_code code, int llr_type, const float *soft, float scale,
                                       int limit, void *llrs, size_t batch, void *cuda_stream);

/* ---------------------------------------------------------------------------
 * Harness kernels (not in the reference's library; they replace the per-trial
 * set-up of its Monte-Carlo driver, perftest/src/main.rs:9-28, on the device).
 * Random numbers are counter-based: Philox4x32-10 keyed by `seed`, counter =
 * (frame index, word index, stream), so frame `first_frame + f` gets the same
 * bits whichever GPU, chunk or call produces it.
 *   random_data : data[B][k/8], 16 bytes per Philox call (stream 1), words
 *                 stored little-endian                         (main.rs:10)
 *   awgn        : y = (1 - 2 bit) + sigma * z, z ~ N(0,1) by Box-Muller on the
 *                 four words of counter (frame, i/4, stream 2) for variables
 *                 i..i+3                                       (main.rs:13-18)
 *                 out_type LABRADOR_LDPC_LLR_F32: out = y * scale
 *                 out_type ..._I8 / ..._I16: out = clamp(rint(y * scale), +-limit)
 *   count_errors: bit_errors[f] = popcount(decoded[f][..k/8] ^ data[f])
 *                 with decoded[B][output_len]                  (main.rs:23-28)
 * ------------------------------------------------------------------------- */
int labrador_ldpc_random_data_batch(enum labrador_ldpc_code code, uint64_t seed, uint64_t first_frame,
                                    uint8_t *data, size_t batch);
int labrador_ldpc_awgn_batch(enum labrador_ldpc_code code, int out_type, const uint8_t *codewords, float sigma,
                             float scale, int limit, uint64_t seed, uint64_t first_frame, void *out, size_t batch);
int labrador_ldpc_count_errors_batch(enum labrador_ldpc_code code, const uint8_t *decoded, const uint8_t *data,
                                     uint32_t *bit_errors, size_t batch);
int labrador_ldpc_random_data_batch_async(enum labrador_ldpc_code code, uint64_t seed, uint64_t first_frame,
                                          uint8_t *data, size_t batch, void *cuda_stream);
int labrador_ldpc_awgn_batch_async(enum labrador_ldpc_code code, int out_type, const uint8_t *codewords, float sigma,
                                   float scale, int limit, uint64_t seed, uint64_t first_frame, void *out,
                                   size_t batch, void *cuda_stream);
int labrador_ldpc_count_errors_batch_async(enum labrador_ldpc_code code, const uint8_t *decoded, const uint8_t *data,
                                           uint32_t *bit_errors, size_t batch, void *cuda_stream);

/* Copy-only control for the benchmark harness: moves the arrays of a host-pointer labrador_ldpc_decode_ms_*_batch
 * call through the same chunked pipeline (H2D of llrs, D2H of output / success / iters_run) without launching a
 * kernel; the result arrays receive unspecified bytes.  Host pointers only. */
int labrador_ldpc_copy_control_batch(enum labrador_ldpc_code code, int llr_type, const void *llrs, uint8_t *output,
                                     size_t batch, uint8_t *success, uint32_t *iters_run);

/* Introspection used by the tests and the benchmark harness. */
/* Number of kernels this library has launched since load (all devices). */
unsigned long long labrador_ldpc_kernel_launch_count(void);
/* Name of the kernel variant decode_ms would use for (code, llr_type), e.g. "ms_generic<i8>". */
const char *labrador_ldpc_decode_ms_kernel_name(enum labrador_ldpc_code code, int llr_type);
/* CRC-32 of the expanded edge table of `code` in reference order (src/codes/mod.rs:508-535). */
uint32_t labrador_ldpc_edge_table_crc(enum labrador_ldpc_code code);
/* Host-only model of the encoders (no GPU needed; used by the CPU tests).  The batched encoders do not multiply by
 * the compact generator as EncodeInto::encode_parity does (src/encoder.rs:42-82): TM codes are encoded through the
 * sparse parity-check matrix and one derived M x M inverse, TC codes through a table of per-byte parity
 * contributions.  This call computes the (n-k)/8 parity bytes of one data block (k/8 bytes) both ways on the host:
 * parity_tables from the very tables the kernels use, parity_generator by the reference's algorithm.
 * Returns 0, or a negative error if a table could not be derived or its two forms disagree.  For the k = 16384 codes,
 * which have no generator, it returns 1 and writes parity_tables only (parity_generator may be NULL). */
int labrador_ldpc_host_encode_model(enum labrador_ldpc_code code, const uint8_t *data, uint8_t *parity_tables,
                                    uint8_t *parity_generator);

#ifdef __cplusplus
}
#endif
#endif /* LABRADOR_LDPC_B200_CAPI */
