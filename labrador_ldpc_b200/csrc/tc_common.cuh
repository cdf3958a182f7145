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
