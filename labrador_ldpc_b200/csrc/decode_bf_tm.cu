// Bit-packed bit-flipping decoder for the TM codes: one codeword per group of min(32, M/32) lanes.
//
// Replaces LDPCCode::decode_bf (reference src/decoder.rs:243-301) and the erasure pre-pass
// decode_erasures (src/decoder.rs:144-223) for TM1280 ... TM8192.  Everything those two functions
// compute is an XOR parity, a small count or a max, so the whole decoder runs on 32 variables /
// checks per machine word:
//   * hard decisions and check parities are bit-packed (bit i of word w = element 32w + i);
//   * the parity of 32 consecutive checks of a block row is the XOR, over the row's blocks, of a
//     32-bit window of the variable bits -- an aligned word for identity blocks, a funnel shift of
//     two neighbouring words for pi_k blocks (quarter q -> (theta+q) mod 4, offset x -> (phi_q + x) mod Q,
//     reference src/codes/mod.rs:312-322), and the other way round for the violated-check counts;
//   * every word is stored twice, in lo[] at its own index and in hi[] at its cyclic predecessor's index
//     inside the quarter, so a window is lo[i] : hi[i] with ONE address; where the window starts (word
//     index, bit shift) depends only on the block and on the lane's word, and is kept in registers for
//     the whole launch; all shared accesses of a warp are bank-conflict free;
//   * the per-variable violation counts (<= 6) come out of a carry-save adder tree as three bit planes;
//     "flip every variable whose count equals the maximum" (src/decoder.rs:276-296) needs only the set
//     of counts that occur, OR-accumulated per count and reduced over the codeword's lanes.
// Codes with fewer than 32 words per block column put 32 / (M/32) codewords in one warp, so every lane
// owns a word; a codeword that has finished idles (no flips) until the rest of its warp is done.
// Erasure pre-pass: as written in the reference it always runs exactly one pass when max_iters >= 1
// and contributes 0 iterations (see decode_bf.cu).  In every TM prototype only the row-2 checks have
// exactly ONE punctured neighbour (through the identity block in the punctured column), so the single
// vote a punctured bit receives is the parity of "its" row-2 check over the transmitted bits; rows 0
// and 1 have two / three punctured neighbours and never vote.  (static_assert'ed below.)
#include <cuda_runtime.h>

#include "bf_common.cuh"
#include "runtime.h"
#include "tm_common.cuh"

namespace ldpc {
using namespace tm;

namespace {

constexpr int kBfWarps = 8;
constexpr int kBfClaim = 8;      // most codeword groups a warp claims per atomic (one hot counter for > 1e9 codewords/s)

template <class P> __host__ __device__ constexpr int blocks_in_row_col(int r, int c) {
    int n = 0;
    for (int b = 0; b < P::NB; b++) n += (P::blk(b).row == r && P::blk(b).col == c);
    return n;
}

// shared-memory words per codeword: lo[] + hi[], padded so the codewords of one warp start LPC banks apart
template <class P, int M> __host__ __device__ constexpr int bf_cw_stride() {
    constexpr int MW = M / 32, LPC = MW < 32 ? MW : 32;
    constexpr int words = 2 * (P::NCOL + P::NROW) * MW;
    return LPC == 32 ? words : ((words + 31) / 32) * 32 + LPC;
}

// does any permutation block read this block column (row parities) / block row (violation counts)?
template <class P> __host__ __device__ constexpr bool col_has_p(int c) {
    for (int b = 0; b < P::NB; b++) if (P::blk(b).col == c && P::blk(b).isp) return true;
    return false;
}
template <class P> __host__ __device__ constexpr bool row_has_p(int r) {
    for (int b = 0; b < P::NB; b++) if (P::blk(b).row == r && P::blk(b).isp) return true;
    return false;
}

__device__ __forceinline__ uint32_t load_be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

template <int RATE, int M>
__global__ void __launch_bounds__(32 * kBfWarps)
decode_bf_tm_kernel(const TmParams prm, const uint8_t *__restrict__ in_all, uint8_t *__restrict__ out_all,
                    unsigned long long batch, unsigned max_iters, uint8_t *__restrict__ success,
                    uint32_t *__restrict__ iters_out, unsigned long long *__restrict__ counter, const int claim_groups) {
    typedef Proto<RATE> P;
    constexpr int NB = P::NB, NCOL = P::NCOL, NROW = P::NROW, CP = NCOL - 1;
    constexpr int NP = count_p<P>(NB);
    constexpr int Q = M / 4, QW = Q / 32;                 // words per quarter
    constexpr int MW = M / 32;                            // words per block row / column
    constexpr int LPC = MW < 32 ? MW : 32;                // lanes per codeword
    constexpr int CWW = 32 / LPC;                         // codewords per warp
    constexpr int WPL = MW / LPC;                         // words per lane per block column
    constexpr int NVW = NCOL * MW, NW = (NCOL - 1) * MW, NCW = NROW * MW;
    constexpr int CWORDS = NVW + NCW;                     // (word, successor) pairs per codeword
    constexpr unsigned kFull = 0xFFFFFFFFu;
    static_assert(Q % 32 == 0 && (QW & (QW - 1)) == 0 && MW % LPC == 0, "quarters are whole words");
    static_assert(blocks_in_row_col<P>(0, CP) == 2 && blocks_in_row_col<P>(1, CP) == 3 &&
                  blocks_in_row_col<P>(2, CP) == 1 && P::blk(NB - 1).row == 2 && P::blk(NB - 1).col == CP &&
                  !P::blk(NB - 1).isp, "only row 2 votes in the erasure pass, through an identity block");

    // per codeword: lo[CWORDS] then hi[CWORDS] (hard decisions first, then parities); the stride between the
    // codewords of one warp is LPC mod 32 words so their lanes fall into disjoint banks
    constexpr int STRIDE = bf_cw_stride<P, M>();
    extern __shared__ __align__(16) uint32_t smem_bf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LPC, wl = lane % LPC;
    uint32_t *lo = smem_bf + (warp * CWW + grp) * STRIDE;
    uint32_t *hi = lo + CWORDS;

    // windows of this lane's words: (word index << 5) | bit shift.
    // direction 0 = check word -> variable bits, direction 1 = variable word -> check parities
    uint32_t win[2][NP > 0 ? NP : 1][WPL];
#pragma unroll
    for (int wi = 0; wi < WPL; wi++) {
        const int e0 = (wl + wi * LPC) * 32, qa = e0 / Q, off = e0 % Q;
        static_for<0, NB>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            if constexpr (P::blk(b).isp) {
                constexpr int ps = count_p<P>(b);
                const int qv = ((int)prm.theta[b] + qa) & 3;
                const int s0 = ((int)prm.phi[b][qa] + off) & (Q - 1);
                win[0][ps][wi] = (uint32_t)(((P::blk(b).col * MW + qv * QW + (s0 >> 5)) << 5) | (s0 & 31));
                const int q = (qa - (int)prm.theta[b]) & 3;
                const int s1 = (off - (int)prm.phi[b][q]) & (Q - 1);
                win[1][ps][wi] = (uint32_t)(((NVW + P::blk(b).row * MW + q * QW + (s1 >> 5)) << 5) | (s1 & 31));
            }
        });
    }

    // word w of an array starting at index `base`: lo[] at its own place, hi[] at its predecessor's
    auto store_word = [&](int base, int w, uint32_t v) {
        lo[base + w] = v;
        hi[base + ((w & ~(QW - 1)) | ((w - 1) & (QW - 1)))] = v;
    };
    auto window = [&](uint32_t t) {
        const uint32_t i = t >> 5;
        return __funnelshift_r(lo[i], hi[i], t);          // the shift is t mod 32
    };
    // Each lane keeps its own words of the hard decisions (bw) and parities (pw) in registers: identity blocks
    // connect word w with word w, so only what a permutation block reads has to go through shared memory.
    uint32_t bw[NCOL][WPL], pw[NROW][WPL];
    // parity word `w` (0..MW-1) of block row r from the current bits
    auto row_parity = [&](auto ri, int wi) {
        constexpr int r = decltype(ri)::value;
        uint32_t x = 0;
        static_for<0, NB>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            if constexpr (P::blk(b).row == r) {
                if constexpr (P::blk(b).isp) x ^= window(win[0][count_p<P>(b)][wi]);
                else x ^= bw[P::blk(b).col][wi];
            }
        });
        return x;
    };

    for (;;) {
        unsigned long long claim = 0;
        if (lane == 0) claim = atomicAdd(counter, (unsigned long long)(CWW * claim_groups));
        claim = __shfl_sync(kFull, claim, 0);
        if (claim >= batch) break;
        for (int ci = 0; ci < claim_groups; ci++) {
            const unsigned long long first = claim + (unsigned long long)(ci * CWW);
            if (first >= batch) break;
            const unsigned long long frame = first + grp;
            const bool live = frame < batch;              // the last warp's trailing groups may not exist
            const uint8_t *in = in_all + (live ? frame : first) * (unsigned long long)(NW * 4);
            const bool in_aligned = (reinterpret_cast<uintptr_t>(in_all) & 3u) == 0;
            // output[..n/8] = input (:251); punctured bits start as zero (:167)
#pragma unroll
            for (int wi = 0; wi < WPL; wi++) {
                const int w = wl + wi * LPC;
#pragma unroll
                for (int c = 0; c < NCOL; c++) {
                    uint32_t v = 0;
                    if (c < NCOL - 1) {
                        const int iw = c * MW + w;
                        v = in_aligned ? __byte_perm(reinterpret_cast<const uint32_t *>(in)[iw], 0, 0x0123)
                                       : load_be32(in + 4 * iw);
                        v = __brev(v);
                    }
                    bw[c][wi] = v;
                    if (col_has_p<P>(c)) store_word(c * MW, w, v);
                }
            }
            __syncwarp();

            if (max_iters > 0) {
                // erasure pass: punctured bit j <- parity of row-2 check j over the transmitted bits (:177-213)
                uint32_t e[WPL];
#pragma unroll
                for (int wi = 0; wi < WPL; wi++) e[wi] = row_parity(std::integral_constant<int, 2>{}, wi);
                __syncwarp();
#pragma unroll
                for (int wi = 0; wi < WPL; wi++) { bw[CP][wi] = e[wi]; store_word(CP * MW, wl + wi * LPC, e[wi]); }
                __syncwarp();
            }

            unsigned iters_run = max_iters;
            bool ok = false, done = !live;
            for (unsigned iter = 0; iter < max_iters; iter++) {
                // parity of every check (:269-273)
                static_for<0, NROW>([&](auto ri) {
                    constexpr int r = decltype(ri)::value;
#pragma unroll
                    for (int wi = 0; wi < WPL; wi++) {
                        pw[r][wi] = row_parity(ri, wi);
                        if constexpr (row_has_p<P>(r)) store_word(NVW + r * MW, wl + wi * LPC, pw[r][wi]);
                    }
                });
                __syncwarp();
                // violated-check count of every variable as bit planes (:276-286), and which counts occur
                uint32_t c0[NCOL][WPL], c1[NCOL][WPL], c2[NCOL][WPL];
                uint32_t occurs[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};       // occurs[v] != 0: some variable has count v
                static_for<0, NCOL>([&](auto ci2) {
                    constexpr int c = decltype(ci2)::value;
                    constexpr int DEG = col_degree<P>(c);
#pragma unroll
                    for (int wi = 0; wi < WPL; wi++) {
                        const int w = wl + wi * LPC;
                        uint32_t x[6] = {0u, 0u, 0u, 0u, 0u, 0u};
                        static_for<0, NB>([&](auto bi) {
                            constexpr int b = decltype(bi)::value;
                            if constexpr (P::blk(b).col == c) {
                                if constexpr (P::blk(b).isp) x[pos_in_col<P>(b)] = window(win[1][count_p<P>(b)][wi]);
                                else x[pos_in_col<P>(b)] = pw[P::blk(b).row][wi];
                            }
                        });
                        uint32_t a0, a1, a2;
                        count_planes<DEG>(x, a0, a1, a2);
                        c0[c][wi] = a0; c1[c][wi] = a1; c2[c][wi] = a2;
#pragma unroll
                        for (int v = 1; v <= DEG; v++)
                            occurs[v] |= ((v & 1) ? a0 : ~a0) & ((v & 2) ? a1 : ~a1) & ((v & 4) ? a2 : ~a2);
                    }
                });
                uint32_t present = 1u;
#pragma unroll
                for (int v = 1; v <= 6; v++) present |= occurs[v] ? (1u << v) : 0u;
#pragma unroll
                for (int d = 1; d < LPC; d <<= 1) present |= __shfl_xor_sync(kFull, present, d);
                const int max_viol = 31 - __clz((int)present);
                if (max_viol == 0 && !done) { ok = true; iters_run = iter; done = true; }   // :288-289
                if (__all_sync(kFull, done)) break;
                // flip every variable whose count equals the maximum (:292-296); finished codewords stay as they are
                if (!done) {
                    const uint32_t m0 = (max_viol & 1) ? kFull : 0u, m1 = (max_viol & 2) ? kFull : 0u,
                                   m2 = (max_viol & 4) ? kFull : 0u;
                    static_for<0, NCOL>([&](auto ci2) {
                        constexpr int c = decltype(ci2)::value;
#pragma unroll
                        for (int wi = 0; wi < WPL; wi++) {
                            bw[c][wi] ^= ~((c0[c][wi] ^ m0) | (c1[c][wi] ^ m1) | (c2[c][wi] ^ m2));
                            if constexpr (col_has_p<P>(c)) store_word(c * MW, wl + wi * LPC, bw[c][wi]);
                        }
                    });
                }
                __syncwarp();
            }

            if (live) {
                uint8_t *out = out_all + frame * (unsigned long long)(NVW * 4);
                const bool aligned = (reinterpret_cast<uintptr_t>(out_all) & 3u) == 0;
#pragma unroll
                for (int wi = 0; wi < WPL; wi++) {
#pragma unroll
                    for (int c = 0; c < NCOL; c++) {
                        const int w = c * MW + wl + wi * LPC;
                        const uint32_t rev = __brev(bw[c][wi]);
                        if (aligned) {
                            reinterpret_cast<uint32_t *>(out)[w] = __byte_perm(rev, 0, 0x0123);
                        } else {
                            out[4 * w + 0] = (uint8_t)(rev >> 24); out[4 * w + 1] = (uint8_t)(rev >> 16);
                            out[4 * w + 2] = (uint8_t)(rev >> 8);  out[4 * w + 3] = (uint8_t)rev;
                        }
                    }
                }
                if (wl == 0) {
                    if (success) success[frame] = ok ? 1 : 0;
                    if (iters_out) iters_out[frame] = iters_run;
                }
            }
            __syncwarp();
        }
    }
}

template <int RATE, int M>
cudaError_t launch_bf_tm(DeviceCtx &ctx, const CodeInfo &c, const uint8_t *input, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    typedef Proto<RATE> P;
    const TmParams prm = make_params<RATE>(c);
    constexpr int MW = M / 32, CWW = MW < 32 ? 32 / MW : 1;            // codewords per warp
    const size_t smem = (size_t)kBfWarps * CWW * bf_cw_stride<P, M>() * sizeof(uint32_t);
    auto kern = decode_bf_tm_kernel<RATE, M>;
    static bool configured[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[ctx.device] = true;
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kBfWarps, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm;
    // large batches claim several groups per atomic; small ones keep one group per claim so every SM gets work
    const unsigned long long groups = (batch + CWW - 1) / CWW;
    unsigned long long claim = groups / (grid * kBfWarps * 4);
    claim = claim < 1 ? 1 : (claim > kBfClaim ? kBfClaim : claim);
    const unsigned long long per_cta = (unsigned long long)kBfWarps * claim;
    const unsigned long long need = (groups + per_cta - 1) / per_cta;
    if (grid > need) grid = need;
    WorkCounter wc(ctx, stream);
    if (wc.error() != cudaSuccess) return wc.error();
    unsigned long long *counter = wc.ptr();
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)grid, 32 * kBfWarps, smem, stream>>>(prm, input, output, (unsigned long long)batch, mi, success,
                                                          iters, counter, (int)claim);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

// Returns true (and launches) for the TM codes.
bool launch_decode_bf_tm(DeviceCtx &ctx, int code, const uint8_t *input, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err) {
    if (code < 3 || code >= kNumCodes) return false;
    const CodeInfo &c = *code_info(code);
    switch (code) {
        // the k = 16384 codes: same prototypes, M = 2048 / 4096 / 8192 (two / four / eight words per lane)
        case 9: if (!structure_matches<2>(c) || c.m != 2048) return false;
                *err = launch_bf_tm<2, 2048>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 10: if (!structure_matches<1>(c) || c.m != 4096) return false;
                *err = launch_bf_tm<1, 4096>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 11: if (!structure_matches<0>(c) || c.m != 8192) return false;
                *err = launch_bf_tm<0, 8192>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 3: if (!structure_matches<2>(c) || c.m != 128) return false;
                *err = launch_bf_tm<2, 128>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 4: if (!structure_matches<1>(c) || c.m != 256) return false;
                *err = launch_bf_tm<1, 256>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 5: if (!structure_matches<0>(c) || c.m != 512) return false;
                *err = launch_bf_tm<0, 512>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 6: if (!structure_matches<2>(c) || c.m != 512) return false;
                *err = launch_bf_tm<2, 512>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 7: if (!structure_matches<1>(c) || c.m != 1024) return false;
                *err = launch_bf_tm<1, 1024>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 8: if (!structure_matches<0>(c) || c.m != 2048) return false;
                *err = launch_bf_tm<0, 2048>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        default: return false;
    }
}

}  // namespace ldpc
