// Shared pieces of the TC-code kernels (decode_ms_tc.cu, decode_bf_tc.cu): block positions of the 4 x 8
// prototype in the reference iterator's order (reference src/codes/compact_parity_checks.rs:21-78,
// src/codes/mod.rs:295-361) and the run-time check that a code's expanded tables match them.
#pragma once
#include <cstdint>
#include <type_traits>

#include "runtime.h"

namespace ldpc {
namespace {

struct TcBlk { int row, col; };
// block positions shared by the three TC prototypes, in the reference iterator's order
__host__ __device__ constexpr TcBlk tc_blk(int b) {
    constexpr TcBlk t[32] = {{0, 0}, {0, 0}, {0, 1}, {0, 2}, {0, 3}, {0, 5}, {0, 6}, {0, 7},
                             {1, 0}, {1, 1}, {1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 6}, {1, 7},
                             {2, 0}, {2, 1}, {2, 2}, {2, 2}, {2, 3}, {2, 4}, {2, 5}, {2, 7},
                             {3, 0}, {3, 1}, {3, 2}, {3, 3}, {3, 3}, {3, 4}, {3, 5}, {3, 6}};
    return t[b];
}
__host__ __device__ constexpr int tc_pos_in_col(int b) {
    int c = 0;
    for (int i = 0; i < b; i++) c += tc_blk(i).col == tc_blk(b).col;
    return c;
}
__host__ __device__ constexpr int tc_pos_in_row(int b) { return b % 8; }

template <int I, int N, class F> __device__ __forceinline__ void tc_static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        tc_static_for<I + 1, N>(f);
    }
}

struct TcParams { uint8_t shift[32]; };

// Rotations of the 32 blocks (reference src/codes/compact_parity_checks.rs:21-78), compile-time so that the index
// arithmetic of equal rotations is shared and folds into the load / store offsets; the launcher checks them against
// the run-time expansion of the prototype tables before a kernel that uses them is launched.
template <int M> __host__ __device__ constexpr int tc_const_shift(int b) {
    constexpr int t16[32] = {0, 7, 2, 14, 6, 0, 13, 0, 6, 0, 15, 0, 1, 0, 0, 7, 4, 1, 0, 15, 14, 11, 0, 3, 0, 1, 9, 0, 13, 14, 1, 0};
    constexpr int t32[32] = {0, 31, 15, 25, 0, 20, 12, 0, 28, 0, 30, 29, 24, 0, 1, 20, 8, 0, 0, 28, 1, 29, 0, 21, 18, 30, 0, 0, 30, 25, 26, 0};
    constexpr int t64[32] = {0, 63, 30, 50, 25, 43, 62, 0, 56, 0, 61, 50, 23, 0, 37, 26, 16, 0, 0, 55, 27, 56, 0, 43, 35, 56, 62, 0, 11, 58, 3, 0};
    return M == 16 ? t16[b] : (M == 32 ? t32[b] : t64[b]);
}
template <int M> inline bool tc_const_shifts_match(const CodeInfo &c) {
    for (int b = 0; b < 32; b++)
        if (c.blocks[b].shift != tc_const_shift<M>(b)) return false;
    return true;
}

inline bool tc_structure_matches(const CodeInfo &c) {
    if (c.n_blocks != 32 || c.rows != 4 || c.cols != 8 || c.p != 0) return false;
    for (int b = 0; b < 32; b++) {
        const Block &blk = c.blocks[b];
        if (blk.kind != kIdentity || blk.row != tc_blk(b).row || blk.col != tc_blk(b).col) return false;
        if (blk.edge_offset != b * c.m || blk.shift < 0 || blk.shift >= c.m) return false;
    }
    return true;
}

}  // namespace
}  // namespace ldpc
