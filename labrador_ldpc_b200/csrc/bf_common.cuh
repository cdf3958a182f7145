// Bit-plane arithmetic shared by the bit-packed bit-flipping kernels (decode_bf_tm.cu, decode_bf_tc.cu).
#pragma once
#include <cstdint>

namespace ldpc {

// sum of up to three / six one-bit planes as bit planes a0 (1), a1 (2), a2 (4): carry-save adders
template <class W> __device__ __forceinline__ void full_add(W x, W y, W z, W &s, W &c) {
    s = x ^ y ^ z;
    c = (x & y) | (z & (x | y));
}
template <int DEG, class W>
__device__ __forceinline__ void count_planes(const W (&x)[6], W &a0, W &a1, W &a2) {
    static_assert(DEG >= 1 && DEG <= 6, "variable degrees of the CCSDS prototypes");
    if constexpr (DEG <= 3) {
        full_add<W>(x[0], DEG > 1 ? x[1] : (W)0, DEG > 2 ? x[2] : (W)0, a0, a1);
        a2 = 0;
    } else {
        W s1, c1, s2, c2;
        full_add<W>(x[0], x[1], x[2], s1, c1);
        full_add<W>(x[3], DEG > 4 ? x[4] : (W)0, DEG > 5 ? x[5] : (W)0, s2, c2);
        a0 = s1 ^ s2;
        full_add<W>(c1, c2, s1 & s2, a1, a2);
    }
}

}  // namespace ldpc
