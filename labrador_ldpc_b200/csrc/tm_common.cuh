// Shared pieces of the specialised TM-code min-sum kernels (decode_ms_tm.cu: packed i8 lanes,
// decode_ms_tm_wide.cu: one element per lane for i16/i32/f32 and for TM1280): compile-time block lists of
// the three TM prototypes in the reference iterator's order, and the launch plumbing.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "runtime.h"

namespace ldpc {
namespace tm {

struct Blk { int row, col, isp; };

template <int RATE> struct Proto;

// Block lists in the reference iterator's order; checked against
// the run-time expansion of the prototype tables before the kernel is ever used.
template <> struct Proto<0> {   // rate 1/2: TM2048 (M=512), TM8192 (M=2048)
    static constexpr int NB = 15, NCOL = 5, NROW = 3;
    __host__ __device__ static constexpr Blk blk(int b) {
        constexpr Blk t[NB] = {{0, 2, 0}, {0, 4, 0}, {0, 4, 1}, {1, 0, 0}, {1, 1, 0}, {1, 3, 0}, {1, 4, 1}, {1, 4, 1},
                               {1, 4, 1}, {2, 0, 0}, {2, 1, 1}, {2, 1, 1}, {2, 3, 1}, {2, 3, 1}, {2, 4, 0}};
        return t[b];
    }
};
template <> struct Proto<1> {   // rate 2/3: TM1536 (M=256), TM6144 (M=1024)
    static constexpr int NB = 23, NCOL = 7, NROW = 3;
    __host__ __device__ static constexpr Blk blk(int b) {
        constexpr Blk t[NB] = {{0, 4, 0}, {0, 6, 0}, {0, 6, 1}, {1, 0, 1}, {1, 0, 1}, {1, 0, 1}, {1, 1, 0}, {1, 2, 0},
                               {1, 3, 0}, {1, 5, 0}, {1, 6, 1}, {1, 6, 1}, {1, 6, 1}, {2, 0, 0}, {2, 1, 1}, {2, 1, 1},
                               {2, 1, 1}, {2, 2, 0}, {2, 3, 1}, {2, 3, 1}, {2, 5, 1}, {2, 5, 1}, {2, 6, 0}};
        return t[b];
    }
};
template <> struct Proto<2> {   // rate 4/5: TM5120 (M=512)  (TM1280, M=128, stays on the generic kernel)
    static constexpr int NB = 39, NCOL = 11, NROW = 3;
    __host__ __device__ static constexpr Blk blk(int b) {
        constexpr Blk t[NB] = {{0, 8, 0},  {0, 10, 0}, {0, 10, 1}, {1, 0, 1},  {1, 0, 1},  {1, 0, 1},  {1, 1, 0},
                               {1, 2, 1},  {1, 2, 1},  {1, 2, 1},  {1, 3, 0},  {1, 4, 1},  {1, 4, 1},  {1, 4, 1},
                               {1, 5, 0},  {1, 6, 0},  {1, 7, 0},  {1, 9, 0},  {1, 10, 1}, {1, 10, 1}, {1, 10, 1},
                               {2, 0, 0},  {2, 1, 1},  {2, 1, 1},  {2, 1, 1},  {2, 2, 0},  {2, 3, 1},  {2, 3, 1},
                               {2, 3, 1},  {2, 4, 0},  {2, 5, 1},  {2, 5, 1},  {2, 5, 1},  {2, 6, 0},  {2, 7, 1},
                               {2, 7, 1},  {2, 9, 1},  {2, 9, 1},  {2, 10, 0}};
        return t[b];
    }
};

template <class P> __host__ __device__ constexpr int count_p(int upto) {
    int c = 0;
    for (int b = 0; b < upto; b++) c += P::blk(b).isp;
    return c;
}
template <class P> __host__ __device__ constexpr int count_i(int upto) { return upto - count_p<P>(upto); }
template <class P> __host__ __device__ constexpr int row_degree(int r) {
    int c = 0;
    for (int b = 0; b < P::NB; b++) c += P::blk(b).row == r;
    return c;
}
template <class P> __host__ __device__ constexpr int col_degree(int col) {
    int c = 0;
    for (int b = 0; b < P::NB; b++) c += P::blk(b).col == col;
    return c;
}
// index (0-based) of block b among the blocks of its row / column
template <class P> __host__ __device__ constexpr int pos_in_row(int b) {
    int c = 0;
    for (int i = 0; i < b; i++) c += P::blk(i).row == P::blk(b).row;
    return c;
}
template <class P> __host__ __device__ constexpr int pos_in_col(int b) {
    int c = 0;
    for (int i = 0; i < b; i++) c += P::blk(i).col == P::blk(b).col;
    return c;
}

template <int I, int N, class F> __device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// theta / phi of every block in block order (only permutation blocks are read)
struct TmParams {
    uint8_t theta[40];
    uint16_t phi[40][4];
};


template <int RATE> bool structure_matches(const CodeInfo &c) {
    typedef Proto<RATE> P;
    if (c.n_blocks != P::NB || c.cols != P::NCOL || c.rows != P::NROW) return false;
    for (int b = 0; b < P::NB; b++) {
        const Block &blk = c.blocks[b];
        const Blk want = P::blk(b);
        if (blk.row != want.row || blk.col != want.col) return false;
        if ((blk.kind == kPermutation) != (want.isp != 0)) return false;
        if (blk.kind == kIdentity && blk.shift != 0) return false;
        if (blk.edge_offset != b * c.m) return false;
    }
    return true;
}

template <int RATE> TmParams make_params(const CodeInfo &c) {
    TmParams prm{};
    for (int b = 0; b < Proto<RATE>::NB; b++) {
        prm.theta[b] = (uint8_t)c.blocks[b].theta;
        for (int j = 0; j < 4; j++) prm.phi[b][j] = (uint16_t)c.blocks[b].phi[j];
    }
    return prm;
}

}  // namespace tm
}  // namespace ldpc
