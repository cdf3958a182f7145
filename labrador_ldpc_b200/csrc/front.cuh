// Fused input front ends of decode_ms.
//
// The reference's callers convert before they decode: `hard_to_llrs` (reference src/decoder.rs:484-493) in the
// examples and the doc tests, and a float -> integer quantiser in front of decode_ms::<i8> (the docs recommend
// leaving headroom, src/decoder.rs:337-340; perftest/src/main.rs:13-18 builds the soft values).  Done as
// separate steps those cost a full extra HBM round trip of n * sizeof(T) bytes per codeword.  Here the
// conversion happens while the decoder loads its channel LLRs, so the input is read from HBM exactly once:
//   kFrontSoftF32: T llr = clamp(rint(soft * scale), -limit, +limit), NaN -> 0 (an erasure); integer T only
//   kFrontHard:    bit-packed hard decisions, MSB first; bit 1 -> -one(), bit 0 -> +one()   (:484-493)
#pragma once
#include <cstdint>

#include "llr_arith.cuh"
#include "runtime.h"

namespace ldpc {

template <int FRONT, class T> struct FrontSrc { typedef T type; };
template <class T> struct FrontSrc<kFrontSoftF32, T> { typedef float type; };
template <class T> struct FrontSrc<kFrontHard, T> { typedef uint8_t type; };

// The exact quantiser (also used by the stand-alone quantise kernel and restated by the oracle's tests).
__device__ __forceinline__ int quantise_soft(float soft, float scale, float limit) {
    float p = __fmul_rn(soft, scale);
    if (p != p) p = 0.0f;
    return (int)fminf(fmaxf(rintf(p), -limit), limit);
}

// Channel LLR of variable i of one frame, in the decoder's compute type.
template <int FRONT, class T>
__device__ __forceinline__ typename Arith<T>::C front_load(const typename FrontSrc<FRONT, T>::type *src, int i,
                                                           float scale, float limit) {
    typedef typename Arith<T>::C C;
    if constexpr (FRONT == kFrontSoftF32) {
        return (C)quantise_soft(src[i], scale, limit);
    } else if constexpr (FRONT == kFrontHard) {
        return ((src[i >> 3] >> (7 - (i & 7))) & 1) ? Arith<T>::neg(Arith<T>::one()) : Arith<T>::one();
    } else {
        return (C)src[i];
    }
}

}  // namespace ldpc
