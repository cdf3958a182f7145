// Bit-packed bit-flipping decoder for the TM codes: one codeword per WARP.
//
// Replaces LDPCCode::decode_bf (reference src/decoder.rs:243-301) and the erasure pre-pass
// decode_erasures (src/decoder.rs:144-223) for TM1280 ... TM8192.  Everything those two functions
// compute is an XOR parity, a small count or a max, so the whole decoder runs on 32 variables /
// checks per machine word:
//   * hard decisions and check parities are bit-packed (bit i of word w = element 32w + i);
//   * the parity of 32 consecutive checks of a block row is the XOR, over the row's blocks, of a
//     32-bit window of the variable bits -- an aligned word for identity blocks, a funnel shift of
//     two words for pi_k blocks (quarter q -> (theta+q) mod 4, offset x -> (phi_q + x) mod Q,
//     reference src/codes/mod.rs:312-322), and the other way round for the violated-check counts;
//   * the per-variable violation counts (<= 6) are bit-sliced 3-bit counters; "flip every variable
//     whose count equals the maximum" (src/decoder.rs:276-296) is a presence mask reduced over the warp.
// Erasure pre-pass: as written in the reference it always runs exactly one pass when max_iters >= 1
// and contributes 0 iterations (see decode_bf.cu).  In every TM prototype only the row-2 checks have
// exactly ONE punctured neighbour (through the identity block in the punctured column), so the single
// vote a punctured bit receives is the parity of "its" row-2 check over the transmitted bits; rows 0
// and 1 have two / three punctured neighbours and never vote.  (static_assert'ed below.)
#include <cuda_runtime.h>

#include "runtime.h"
#include "tm_common.cuh"

namespace ldpc {
using namespace tm;

namespace {

constexpr int kBfWarps = 8;

template <class P> __host__ __device__ constexpr int blocks_in_row_col(int r, int c) {
    int n = 0;
    for (int b = 0; b < P::NB; b++) n += (P::blk(b).row == r && P::blk(b).col == c);
    return n;
}

__device__ __forceinline__ uint32_t load_be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

template <int RATE, int M>
__global__ void __launch_bounds__(32 * kBfWarps)
decode_bf_tm_kernel(const TmParams prm, const uint8_t *__restrict__ in_all, uint8_t *__restrict__ out_all,
                    unsigned long long batch, unsigned max_iters, uint8_t *__restrict__ success,
                    uint32_t *__restrict__ iters_out, unsigned long long *__restrict__ counter) {
    typedef Proto<RATE> P;
    constexpr int NB = P::NB, NCOL = P::NCOL, NROW = P::NROW, CP = NCOL - 1;
    constexpr int Q = M / 4, QW = Q / 32;                 // words per quarter
    constexpr int MW = M / 32;                            // words per block row / column
    constexpr int NVW = NCOL * MW, NW = (NCOL - 1) * MW, NCW = NROW * MW;
    constexpr int WPL = (MW + 31) / 32;                   // words per lane per block column
    constexpr unsigned kFull = 0xFFFFFFFFu;
    static_assert(Q % 32 == 0, "quarters are whole words");
    static_assert(blocks_in_row_col<P>(0, CP) == 2 && blocks_in_row_col<P>(1, CP) == 3 &&
                  blocks_in_row_col<P>(2, CP) == 1 && P::blk(NB - 1).row == 2 && P::blk(NB - 1).col == CP &&
                  !P::blk(NB - 1).isp, "only row 2 votes in the erasure pass, through an identity block");

    extern __shared__ __align__(16) uint32_t smem_bf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *bits = smem_bf + warp * (NVW + NCW);        // [NVW] hard decisions
    uint32_t *par = bits + NVW;                            // [NCW] check parities

    // parity word `w` (0..MW-1) of block row r from the current bits
    auto row_parity = [&](auto ri, int w) {
        constexpr int r = decltype(ri)::value;
        const int i0 = w * 32, q = i0 / Q, iq0 = i0 % Q;
        uint32_t x = 0;
        static_for<0, NB>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            if constexpr (P::blk(b).row == r) {
                constexpr int col = P::blk(b).col;
                if constexpr (P::blk(b).isp) {
                    const int qv = ((int)prm.theta[b] + q) & 3;
                    const int s = ((int)prm.phi[b][q] + iq0) & (Q - 1);
                    const int base = col * MW + qv * QW;
                    const int w0 = s >> 5, w1 = (w0 + 1) & (QW - 1);
                    x ^= __funnelshift_r(bits[base + w0], bits[base + w1], s & 31);
                } else {
                    x ^= bits[col * MW + w];
                }
            }
        });
        return x;
    };

    for (;;) {
        unsigned long long frame = 0;
        if (lane == 0) frame = atomicAdd(counter, 1ull);
        frame = __shfl_sync(kFull, frame, 0);
        if (frame >= batch) break;
        const uint8_t *in = in_all + frame * (unsigned long long)(NW * 4);
        // output[..n/8] = input (:251); punctured bits start as zero (:167)
        for (int w = lane; w < NVW; w += 32) bits[w] = w < NW ? __brev(load_be32(in + 4 * w)) : 0u;
        __syncwarp();

        if (max_iters > 0) {
            // erasure pass: punctured bit j <- parity of row-2 check j over the transmitted bits (:177-213)
            for (int w = lane; w < MW; w += 32) par[w] = row_parity(std::integral_constant<int, 2>{}, w);
            __syncwarp();
            for (int w = lane; w < MW; w += 32) bits[CP * MW + w] = par[w];
            __syncwarp();
        }

        unsigned iters_run = max_iters;
        bool ok = false;
        for (unsigned iter = 0; iter < max_iters; iter++) {
            // parity of every check (:269-273)
            static_for<0, NROW>([&](auto ri) {
                constexpr int r = decltype(ri)::value;
                for (int w = lane; w < MW; w += 32) par[r * MW + w] = row_parity(ri, w);
            });
            __syncwarp();
            // violated-check count of every variable, bit-sliced (:276-286)
            uint32_t c0[NCOL][WPL], c1[NCOL][WPL], c2[NCOL][WPL];
            uint32_t present = 0;                       // bit v set: some variable has count v
            static_for<0, NCOL>([&](auto ci) {
                constexpr int c = decltype(ci)::value;
#pragma unroll
                for (int wi = 0; wi < WPL; wi++) {
                    const int w = lane + wi * 32;
                    uint32_t a0 = 0, a1 = 0, a2 = 0;
                    if (w < MW) {
                        const int j0 = w * 32, qv = j0 / Q, jq0 = j0 % Q;
                        static_for<0, NB>([&](auto bi) {
                            constexpr int b = decltype(bi)::value;
                            if constexpr (P::blk(b).col == c) {
                                constexpr int r = P::blk(b).row;
                                uint32_t x;
                                if constexpr (P::blk(b).isp) {
                                    const int q = (qv - (int)prm.theta[b]) & 3;
                                    const int s = (jq0 - (int)prm.phi[b][q]) & (Q - 1);
                                    const int base = r * MW + q * QW;
                                    const int w0 = s >> 5, w1 = (w0 + 1) & (QW - 1);
                                    x = __funnelshift_r(par[base + w0], par[base + w1], s & 31);
                                } else {
                                    x = par[r * MW + w];
                                }
                                const uint32_t t0 = a0 & x;
                                a0 ^= x;
                                const uint32_t t1 = a1 & t0;
                                a1 ^= t0;
                                a2 ^= t1;
                            }
                        });
                        // which counts occur in this word
                        const uint32_t n0 = ~a0, n1 = ~a1, n2 = ~a2;
                        present |= ((n2 & n1 & n0) ? 1u : 0u) | ((n2 & n1 & a0) ? 2u : 0u) | ((n2 & a1 & n0) ? 4u : 0u) |
                                   ((n2 & a1 & a0) ? 8u : 0u) | ((a2 & n1 & n0) ? 16u : 0u) | ((a2 & n1 & a0) ? 32u : 0u) |
                                   ((a2 & a1 & n0) ? 64u : 0u) | ((a2 & a1 & a0) ? 128u : 0u);
                    }
                    c0[c][wi] = a0; c1[c][wi] = a1; c2[c][wi] = a2;
                }
            });
            present = __reduce_or_sync(kFull, present);
            const int max_viol = 31 - __clz((int)present);
            if (max_viol == 0) { ok = true; iters_run = iter; break; }          // :288-289
            // flip every variable whose count equals the maximum (:292-296)
            const uint32_t m0 = (max_viol & 1) ? kFull : 0u, m1 = (max_viol & 2) ? kFull : 0u, m2 = (max_viol & 4) ? kFull : 0u;
            static_for<0, NCOL>([&](auto ci) {
                constexpr int c = decltype(ci)::value;
#pragma unroll
                for (int wi = 0; wi < WPL; wi++) {
                    const int w = lane + wi * 32;
                    if (w < MW) bits[c * MW + w] ^= ~((c0[c][wi] ^ m0) | (c1[c][wi] ^ m1) | (c2[c][wi] ^ m2));
                }
            });
            __syncwarp();
        }

        uint8_t *out = out_all + frame * (unsigned long long)(NVW * 4);
        const bool aligned = (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
        for (int w = lane; w < NVW; w += 32) {
            const uint32_t rev = __brev(bits[w]);
            if (aligned) {
                reinterpret_cast<uint32_t *>(out)[w] = __byte_perm(rev, 0, 0x0123);
            } else {
                out[4 * w + 0] = (uint8_t)(rev >> 24); out[4 * w + 1] = (uint8_t)(rev >> 16);
                out[4 * w + 2] = (uint8_t)(rev >> 8);  out[4 * w + 3] = (uint8_t)rev;
            }
        }
        if (lane == 0) {
            if (success) success[frame] = ok ? 1 : 0;
            if (iters_out) iters_out[frame] = iters_run;
        }
        __syncwarp();
    }
}

template <int RATE, int M>
cudaError_t launch_bf_tm(DeviceCtx &ctx, const CodeInfo &c, const uint8_t *input, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    typedef Proto<RATE> P;
    const TmParams prm = make_params<RATE>(c);
    const size_t smem = (size_t)kBfWarps * (P::NCOL + P::NROW) * (M / 32) * sizeof(uint32_t);
    auto kern = decode_bf_tm_kernel<RATE, M>;
    static bool configured[16] = {};
    if (!configured[ctx.device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[ctx.device & 15] = true;
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kBfWarps, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm;
    const unsigned long long need = (batch + kBfWarps - 1) / kBfWarps;
    if (grid > need) grid = need;
    unsigned long long *counter = nullptr;
    e = next_counter(ctx.device, stream, &counter);
    if (e != cudaSuccess) return e;
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)grid, 32 * kBfWarps, smem, stream>>>(prm, input, output, (unsigned long long)batch, mi, success,
                                                          iters, counter);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

// Returns true (and launches) for the TM codes.
bool launch_decode_bf_tm(DeviceCtx &ctx, int code, const uint8_t *input, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err) {
    if (code < 3 || code > 8) return false;
    const CodeInfo &c = *code_info(code);
    switch (code) {
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
