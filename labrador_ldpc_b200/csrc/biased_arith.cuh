// Biased integer arithmetic of the scalar-lane min-sum kernels for i8 / i16 LLRs (decode_ms_tm_wide.cu,
// decode_ms_tc.cu): the representation of the packed i8 kernel (decode_ms_tm.cu), one value per 32-bit register.
//   B = 2^(BITS-1), MAXV = B - 1 (the type's maximum)
//   marginal      VA = va + B in [0, 2B-1];  saturating_add(va, u) = relu(min(VA + u, 2B-1)): one VIADDMNMX.RELU
//   var -> check  C = MAXV - clamp(va - u, -MAXV, MAXV) = relu(min((2B-1 - VA) + u, 2 MAXV)), in [0, 2 MAXV]
//                 (-B and -MAXV are interchangeable: v is only used through saturating_abs, sign and == 0)
//   check side    sign(v) = bit BITS-1 of C, |v| = |C - MAXV|, v == 0 <=> C == MAXV;
//                 self-correction (reference src/decoder.rs:422-426): kill = that bit of (C ^ old) & (C ^ (old + 1))
#pragma once
#include <cstdint>

namespace ldpc {

// all four bytes <- the most significant bit of byte BYTE of x (0xFFFFFFFF or 0)
template <int BYTE> __device__ __forceinline__ uint32_t sign_mask_of_byte(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(x), "n"(0x1111 * (8 + BYTE)));
    return r;
}

// u_k = min over the other edges of one check, three-input minima at pair boundaries (see decode_ms_tm.cu)
template <int DC, int MAXDEG>
__device__ __forceinline__ void min_excluding_self_u32(const uint32_t (&a)[MAXDEG], uint32_t (&mu)[MAXDEG]) {
    constexpr int NPAIR = DC / 2;
    constexpr bool ODD = (DC & 1) != 0;
    uint32_t suf[MAXDEG / 2 + 2];
    if constexpr (ODD) suf[NPAIR] = a[DC - 1];
#pragma unroll
    for (int j = NPAIR - 1; j >= 1; j--) {
        if (j == NPAIR - 1 && !ODD) suf[j] = min(a[2 * j], a[2 * j + 1]);
        else suf[j] = __vimin3_u32(a[2 * j], a[2 * j + 1], suf[j + 1]);
    }
    uint32_t pre = 0;
#pragma unroll
    for (int j = 0; j < NPAIR; j++) {
        const bool has_pre = j > 0, has_suf = (j + 1 < NPAIR) || ODD;
        if (has_pre && has_suf) {
            mu[2 * j] = __vimin3_u32(pre, a[2 * j + 1], suf[j + 1]);
            mu[2 * j + 1] = __vimin3_u32(pre, a[2 * j], suf[j + 1]);
        } else if (has_suf) {
            mu[2 * j] = min(a[2 * j + 1], suf[j + 1]);
            mu[2 * j + 1] = min(a[2 * j], suf[j + 1]);
        } else if (has_pre) {
            mu[2 * j] = min(pre, a[2 * j + 1]);
            mu[2 * j + 1] = min(pre, a[2 * j]);
        } else {
            mu[2 * j] = a[2 * j + 1];
            mu[2 * j + 1] = a[2 * j];
        }
        if (j + 1 < NPAIR || ODD) pre = has_pre ? __vimin3_u32(pre, a[2 * j], a[2 * j + 1]) : min(a[2 * j], a[2 * j + 1]);
    }
    if constexpr (ODD) mu[DC - 1] = pre;
}

}  // namespace ldpc
