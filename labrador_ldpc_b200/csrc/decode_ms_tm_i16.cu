// Specialised min-sum decoder for the TM codes with i16 LLRs: 32-bit variable side, packed 16-bit check side.
//
// Replaces LDPCCode::decode_ms::<i16> (reference src/decoder.rs:347-475; DecodeFrom for i16, :52-58) for
// TM1280 ... TM8192; results (decoded bytes, success flag, iteration count) are bit-identical to the reference's.
//
// Same skeleton as the packed i8 kernel (decode_ms_tm.cu: lane pairs (x, x + S) inside every quarter of a block, thread t
// owns word slot t of every prototype column and row, identity-block messages in registers, pi_k-block messages in
// shared memory in check order, next frame staged by a bulk asynchronous copy while the current one is decoded) with the
// biased representation of biased_arith.cuh at 16 bits:  B = 32768, MAXV = 32767,
//   marginal       VA = va + B in [0, 65535]
//   var -> check   C = MAXV - clamp(va - u, -MAXV, MAXV) in [0, 65534]
//   check side     sign = bit 15 of C, |v| = |C - MAXV|, v == 0 <=> C == MAXV,
//                  self-correction (:422-426): kill = bit 15 of (C ^ old) & (C ^ (old + 1)).
// C, |v|, the minima and u = +-min fit 16-bit lanes, so the CHECK side -- three quarters of the ALU-pipe work of the
// scalar-lane kernel (decode_ms_tm_wide.cu) -- runs two edges per instruction, exactly like the i8 kernel.  The saturating
// adds of the VARIABLE side need 17 bits (VIADDMNMX.S16x2 wraps, tools/ubench/sat16.cu), so there each lane of a message
// is sign-extended into its own 32-bit register (one PRMT / SHF), added with VIADDMNMX.S32.RELU, and the two results
// are packed again by one IMAD (FMA pipe).
//
// Exit test (:445-453) in the threads that own the checks, as KNOBS bit 5 of the i8 kernel: row 0 is I(CA) + I(CP) +
// P(CP); the two identity terms are the thread's own marginals, the permuted one arrives through a message-shaped
// array of packed marginals (`hmsg`, written with the address and lane swap of block 2: a 16-bit message has no spare
// bits to carry it).  Rows 1-2 are only tested when row 0 is clean, from ballot-packed hard bits.
#include <cuda_runtime.h>

#include <cstdlib>

#include "biased_arith.cuh"
#include "bulk_copy.cuh"
#include "front.cuh"
#include "runtime.h"
#include "tm_common.cuh"

namespace ldpc {
using namespace tm;

namespace {

constexpr int kMaxDeg16 = 18;

__device__ __forceinline__ uint32_t rot16(uint32_t x, uint32_t sh) { return __funnelshift_l(x, x, sh); }
// bit 15 of each 16-bit lane -> lane mask
__device__ __forceinline__ uint32_t sign15_mask(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(r) : "r"(x));
    return r;
}
// low lane of a packed message, sign-extended
__device__ __forceinline__ int lane0_s32(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0x9910;" : "=r"(r) : "r"(x));
    return (int)r;
}

// minimum over the other edges of one check word, three-input minima at pair boundaries (as decode_ms_tm.cu)
template <int DC>
__device__ __forceinline__ void min_excluding_self16(const uint32_t (&a)[kMaxDeg16], uint32_t (&mu)[kMaxDeg16]) {
    constexpr int NPAIR = DC / 2;
    constexpr bool ODD = (DC & 1) != 0;
    uint32_t suf[kMaxDeg16 / 2 + 2];
    if constexpr (ODD) suf[NPAIR] = a[DC - 1];
#pragma unroll
    for (int j = NPAIR - 1; j >= 1; j--) {
        if (j == NPAIR - 1 && !ODD) suf[j] = __vminu2(a[2 * j], a[2 * j + 1]);
        else suf[j] = __vimin3_u16x2(a[2 * j], a[2 * j + 1], suf[j + 1]);
    }
    uint32_t pre = 0;
#pragma unroll
    for (int j = 0; j < NPAIR; j++) {
        const bool has_pre = j > 0, has_suf = (j + 1 < NPAIR) || ODD;
        if (has_pre && has_suf) {
            mu[2 * j] = __vimin3_u16x2(pre, a[2 * j + 1], suf[j + 1]);
            mu[2 * j + 1] = __vimin3_u16x2(pre, a[2 * j], suf[j + 1]);
        } else if (has_suf) {
            mu[2 * j] = __vminu2(a[2 * j + 1], suf[j + 1]);
            mu[2 * j + 1] = __vminu2(a[2 * j], suf[j + 1]);
        } else if (has_pre) {
            mu[2 * j] = __vminu2(pre, a[2 * j + 1]);
            mu[2 * j + 1] = __vminu2(pre, a[2 * j]);
        } else {
            mu[2 * j] = a[2 * j + 1];
            mu[2 * j + 1] = a[2 * j];
        }
        if (j + 1 < NPAIR || ODD)
            pre = has_pre ? __vimin3_u16x2(pre, a[2 * j], a[2 * j + 1]) : __vminu2(a[2 * j], a[2 * j + 1]);
    }
    if constexpr (ODD) mu[DC - 1] = pre;
}

template <int RATE, int M, int WPT, int MINB, int FRONT>
__global__ void __launch_bounds__(M / 2 / WPT, MINB)
decode_ms_tm_i16_kernel(const TmParams prm, const typename FrontSrc<FRONT, int16_t>::type *__restrict__ llrs_all,
                        uint8_t *__restrict__ out_all, unsigned long long batch, unsigned max_iters,
                        uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out,
                        unsigned long long *__restrict__ counter,
                        const uint32_t one /* == 1: keeps carry-free packing and subtractions on the FMA pipe (IMAD) */,
                        const float fscale, const float flimit) {
    typedef typename FrontSrc<FRONT, int16_t>::type Src;
    typedef Proto<RATE> P;
    constexpr int NB = P::NB, NCOL = P::NCOL, NROW = P::NROW;
    constexpr int NP = count_p<P>(NB), NI = NB - NP;
    constexpr int Q = M / 4, S = Q / 2, NT = M / 2 / WPT;
    constexpr int NV = NCOL * M, N = (NCOL - 1) * M, NC = NROW * M;
    constexpr int HBW = NV / 32, SYW = NC / 32;
    constexpr int CA = P::blk(0).col, CP = NCOL - 1;
    constexpr int PS2 = count_p<P>(2);               // block 2 = P(CP) in row 0
    static_assert(S >= 16 && NT % 32 == 0 && SYW <= NT, "whole warps; one thread per syndrome word");
    static_assert(P::blk(0).row == 0 && !P::blk(0).isp && P::blk(1).row == 0 && P::blk(1).col == CP && !P::blk(1).isp &&
                  P::blk(2).row == 0 && P::blk(2).col == CP && P::blk(2).isp && P::blk(3).row == 1,
                  "row 0 must be I(CA) + I(CP) + P(CP)");

    extern __shared__ __align__(16) uint32_t smem_i16[];
    uint32_t *msg = smem_i16;                        // [NP][M/2] permutation-block messages, check order
    uint32_t *hmsg = msg + NP * (M / 2);             // [M/2] packed marginals of column CP as block 2 permutes them
    uint32_t *hb = hmsg + M / 2;                     // [HBW] packed hard decisions (second stage / output)
    constexpr unsigned FB = FRONT == kFrontSoftF32 ? N * 4 : N * 2;     // input bytes per frame
    static_assert(FB % 16 == 0, "bulk copies move multiples of 16 bytes");
    unsigned char *stage = reinterpret_cast<unsigned char *>(hb + ((HBW + 3) & ~3));   // [2][FB]
    const unsigned char *in_all = reinterpret_cast<const unsigned char *>(llrs_all);
    __shared__ unsigned long long s_frame[2];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31;

    uint32_t paddr[NP > 0 ? NP : 1][WPT], pswp[NP > 0 ? NP : 1][WPT];
    uint32_t hbw[WPT];
#pragma unroll
    for (int wi = 0; wi < WPT; wi++) {
        const int wd = tid + wi * NT;
        const int qv = wd / S, wv = wd % S;
        hbw[wi] = (uint32_t)((qv * Q + wv) >> 5);
        static_for<0, NB>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            if constexpr (P::blk(b).isp) {
                constexpr int ps = count_p<P>(b);
                const int q = (qv - (int)prm.theta[b]) & 3;
                const int phi = prm.phi[b][q];
                const int phi_lo = phi % S, phi_hi = phi / S;
                const int borrow = wv < phi_lo ? 1 : 0;
                const int w = (wv - phi_lo) & (S - 1);
                paddr[ps][wi] = (uint32_t)(ps * (M / 2) + q * S + w);
                pswp[ps][wi] = ((phi_hi ^ borrow) & 1) ? 16u : 0u;
            }
        });
    }
    const uint32_t c65535 = 65535u * one, c65536 = one << 16;

    const bool use_bulk = (reinterpret_cast<uintptr_t>(llrs_all) & 15u) == 0;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init_fence();
        const unsigned long long f0 = atomicAdd(counter, 1ull);
        s_frame[0] = f0;
        if (use_bulk && f0 < batch) bulk_load(stage, in_all + f0 * (unsigned long long)FB, FB, &s_bar[0]);
    }
    __syncthreads();
    unsigned cur = 0, bar_parity = 0;

    for (;;) {
        const unsigned long long frame = s_frame[cur];
        if (frame >= batch) break;
        if (tid == 0) {
            const unsigned long long fn = atomicAdd(counter, 1ull);
            s_frame[cur ^ 1] = fn;
            if (use_bulk && fn < batch)
                bulk_load(stage + (cur ^ 1) * FB, in_all + fn * (unsigned long long)FB, FB, &s_bar[cur ^ 1]);
        }
        const Src *llr;
        if (use_bulk) {
            mbar_wait(&s_bar[cur], (bar_parity >> cur) & 1u);
            bar_parity ^= 1u << cur;
            llr = reinterpret_cast<const Src *>(stage + cur * FB);
        } else {
            llr = reinterpret_cast<const Src *>(in_all + frame * (unsigned long long)FB);
        }

        // ---- per-frame state: everything zero, every call (:368, :374) ----
        uint32_t Lb[NCOL][WPT];               // channel LLR + 32768, two lanes (punctured column: 32768)
        uint32_t idm[NI > 0 ? NI : 1][WPT];   // identity-block messages (u after the check phase, C after the variable phase)
        uint32_t cc[NB][WPT];                 // corrected C of the previous iteration (sign class of the old v)
#pragma unroll
        for (int wi = 0; wi < WPT; wi++) {
            const int wd = tid + wi * NT;
            const int e0 = (wd / S) * Q + (wd % S);
#pragma unroll
            for (int c = 0; c < NCOL; c++) {
                if (c < NCOL - 1) {
                    const int l0 = front_load<FRONT, int16_t>(llr, c * M + e0, fscale, flimit);
                    const int l1 = front_load<FRONT, int16_t>(llr, c * M + e0 + S, fscale, flimit);
                    Lb[c][wi] = (uint32_t)(l0 + 32768) | ((uint32_t)(l1 + 32768) << 16);
                } else {
                    Lb[c][wi] = 0x80008000u;                                      // :383
                }
            }
#pragma unroll
            for (int i = 0; i < NI; i++) idm[i][wi] = 0;
#pragma unroll
            for (int b = 0; b < NB; b++) cc[b][wi] = 0x7fff7fffu;
#pragma unroll
            for (int p = 0; p < NP; p++) msg[p * (M / 2) + wd] = 0;
        }
        for (int i = tid; i < HBW; i += NT) hb[i] = 0;
        __syncthreads();

        unsigned iters_run = max_iters;
        bool ok = false, hb_complete = true;
        uint32_t gat[NCOL][WPT];              // packed marginals of the latest variable phase (bit 15 / 31 set: hard bit 0)
        uint32_t hloc[WPT], bad[WPT];
        auto flush_pack = [&]() {             // hard bits of every column into hb[] (second stage / output)
            if constexpr (S % 32 != 0) {
                for (int i = tid; i < HBW; i += NT) hb[i] = 0;
                __syncthreads();
            }
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    const uint32_t g = gat[c][wi];
                    if constexpr (S % 32 == 0) {
                        const unsigned b0 = __ballot_sync(0xFFFFFFFFu, (g & 0x00008000u) == 0);
                        const unsigned b1 = __ballot_sync(0xFFFFFFFFu, (g & 0x80000000u) == 0);
                        if (lane == 0) {
                            hb[hbw[wi] + c * M / 32] = b0;
                            hb[hbw[wi] + (c * M + S) / 32] = b1;
                        }
                    } else {
                        const int wd = tid + wi * NT;
                        const int e0 = c * M + (wd / S) * Q + (wd % S);
                        if ((g & 0x00008000u) == 0) atomicOr(&hb[e0 >> 5], 1u << (e0 & 31));
                        if ((g & 0x80000000u) == 0) atomicOr(&hb[(e0 + S) >> 5], 1u << ((e0 + S) & 31));
                    }
                });
            }
        };

        for (unsigned iter = 0; iter < max_iters; iter++) {
            // ================= variable phase (:382-411 and :421), 32-bit lanes =================
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    int va0 = (int)(Lb[c][wi] & 0xffffu), va1 = (int)(Lb[c][wi] >> 16);
                    int u0[6], u1[6];
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            uint32_t u;
                            if constexpr (P::blk(b).isp) {
                                constexpr int ps = count_p<P>(b);
                                u = rot16(msg[paddr[ps][wi]], pswp[ps][wi]);
                            } else {
                                u = idm[count_i<P>(b)][wi];
                            }
                            u0[k] = lane0_s32(u);
                            u1[k] = (int)u >> 16;
                            va0 = __viaddmin_s32_relu(va0, u0[k], 65535);            // saturating_add, ascending idx (:408)
                            va1 = __viaddmin_s32_relu(va1, u1[k], 65535);
                        }
                    });
                    const int van0 = (int)(c65535 * one) - va0, van1 = (int)(c65535 * one) - va1;
                    const uint32_t pm = (uint32_t)va1 * c65536 + (uint32_t)va0;      // packed marginals (IMAD, no carry)
                    gat[c][wi] = pm;
                    if constexpr (c == CA) hloc[wi] = pm;
                    if constexpr (c == CP) hloc[wi] ^= pm;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            // C = MAXV - clamp(va - u, -MAXV, MAXV)
                            const int cv0 = __viaddmin_s32_relu(van0, u0[k], 65534);
                            const int cv1 = __viaddmin_s32_relu(van1, u1[k], 65534);
                            const uint32_t cv = (uint32_t)cv1 * c65536 + (uint32_t)cv0;
                            if constexpr (P::blk(b).isp) {
                                constexpr int ps = count_p<P>(b);
                                msg[paddr[ps][wi]] = rot16(cv, pswp[ps][wi]);
                                if constexpr (b == 2) hmsg[paddr[ps][wi] - PS2 * (M / 2)] = rot16(pm, pswp[ps][wi]);
                            } else {
                                idm[count_i<P>(b)][wi] = cv;
                            }
                        }
                    });
                });
            }
            __syncthreads();

            // ================= check phase (:391-405 and :422-447), packed 16-bit lanes =================
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                const int wd = tid + wi * NT;
                static_for<0, NROW>([&](auto ri) {
                    constexpr int r = decltype(ri)::value;
                    constexpr int DC = row_degree<P>(r);
                    uint32_t a[kMaxDeg16], ck[kMaxDeg16], mu[kMaxDeg16];
                    uint32_t sx = 0;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            uint32_t cv;
                            if constexpr (P::blk(b).isp) cv = msg[count_p<P>(b) * (M / 2) + wd];
                            else cv = idm[count_i<P>(b)][wi];
                            if constexpr (b == 2) bad[wi] = ~(hloc[wi] ^ hmsg[wd]) & 0x80008000u;   // three biased sign bits XORed = NOT parity
                            const uint32_t old = cc[b][wi];
                            const uint32_t x = (cv ^ old) & (cv ^ (old + 0x00010001u));   // bit 15: sign flipped and old != 0
                            const uint32_t km = sign15_mask(x);
                            const uint32_t cor = (cv & ~km) | (0x7fff7fffu & km);         // killed -> v = 0
                            cc[b][wi] = cor;
                            ck[k] = cor;
                            const uint32_t d = __vsub2(cor, 0x7fff7fffu);                  // -v per lane, in [-32767, 32767]
                            a[k] = __vmaxs2(d, __vsub2(0u, d));                           // |v|
                            sx ^= cor;                                                     // bit 15: product of signs
                        }
                    });
                    min_excluding_self16<DC>(a, mu);
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            const uint32_t nm = sign15_mask(sx ^ ck[k]);                   // lanes whose u is negative
                            const uint32_t u = __vadd2(mu[k], nm) ^ nm;                    // +-mu, two's complement
                            if constexpr (P::blk(b).isp) msg[count_p<P>(b) * (M / 2) + wd] = u;
                            else idm[count_i<P>(b)][wi] = u;
                        }
                    });
                });
            }
            // ---- parity of the marginals' hard bits (:445-453) ----
            auto syndrome_word = [&](int sw) {
                uint32_t synd = 0;
                const int i0 = (sw * 32) % M, r = (sw * 32) / M;
                const int q = i0 / Q, iq0 = i0 % Q;
                static_for<0, NB>([&](auto bi) {
                    constexpr int b = decltype(bi)::value;
                    if (P::blk(b).row == r) {
                        constexpr int col = P::blk(b).col;
                        if constexpr (P::blk(b).isp) {
                            const int qv = ((int)prm.theta[b] + q) & 3;
                            const int s = ((int)prm.phi[b][q] + iq0) & (Q - 1);
                            const int base = (col * M + qv * Q) >> 5;
                            const int w0 = s >> 5, w1 = (w0 + 1) & (Q / 32 - 1);
                            synd ^= __funnelshift_r(hb[base + w0], hb[base + w1], s & 31);
                        } else {
                            synd ^= hb[(col * M + i0) >> 5];
                        }
                    }
                });
                return synd;
            };
            uint32_t synd = 0;
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) synd |= bad[wi];
            hb_complete = false;
            if (__syncthreads_or(synd != 0) == 0) {
                flush_pack();                    // row 0 is clean: pack every column, test rows 1..NROW-1
                hb_complete = true;
                __syncthreads();
                synd = 0;
                for (int sw = M / 32 + tid; sw < SYW; sw += NT) synd |= syndrome_word(sw);
                if (__syncthreads_or(synd != 0) == 0) {
                    ok = true;
                    iters_run = iter;                                                      // :462
                    break;
                }
            }
        }
        if (!hb_complete) {       // decoding failed: the output is the hard decision of the last marginals (:466-473)
            flush_pack();
            __syncthreads();
        }

        // ---- output: hard decisions of all n+p marginals, MSB first (:455-461, :466-473) ----
        uint8_t *out = out_all + frame * (unsigned long long)(NV / 8);
        const bool aligned = (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
        for (int i = tid; i < HBW; i += NT) {
            const uint32_t rev = __brev(hb[i]);
            if (aligned) {
                reinterpret_cast<uint32_t *>(out)[i] = __byte_perm(rev, 0, 0x0123);
            } else {
                out[4 * i + 0] = (uint8_t)(rev >> 24); out[4 * i + 1] = (uint8_t)(rev >> 16);
                out[4 * i + 2] = (uint8_t)(rev >> 8);  out[4 * i + 3] = (uint8_t)rev;
            }
        }
        if (tid == 0) {
            if (success) success[frame] = ok ? 1 : 0;
            if (iters_out) iters_out[frame] = iters_run;
        }
        __syncthreads();   // hb / msg / stage / s_frame are reused by the next frame
        cur ^= 1;
    }
}

template <int RATE, int M, int WPT, int MINB, int FRONT>
cudaError_t launch_i16(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                       size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, const Front &front) {
    typedef Proto<RATE> P;
    constexpr int NP = count_p<P>(P::NB);
    constexpr int NT = M / 2 / WPT;
    const TmParams prm = make_params<RATE>(c);
    const size_t smem = ((size_t)(NP + 1) * (M / 2) + (((size_t)P::NCOL * M / 32 + 3) & ~(size_t)3)) * sizeof(uint32_t) +
                        2 * front_frame_bytes(front, (P::NCOL - 1) * M, kI16);
    auto kern = decode_ms_tm_i16_kernel<RATE, M, WPT, MINB, FRONT>;
    static bool configured[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[ctx.device] = true;
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm;
    if (grid > batch) grid = batch;
    WorkCounter wc(ctx, stream);
    if (wc.error() != cudaSuccess) return wc.error();
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)grid, NT, smem, stream>>>(prm, static_cast<const typename FrontSrc<FRONT, int16_t>::type *>(llrs), output,
                                               (unsigned long long)batch, mi, success, iters, wc.ptr(), 1u, front.scale,
                                               front.limit);
    count_launch();
    return cudaGetLastError();
}

template <int RATE, int M, int WPT, int MINB>
cudaError_t launch_i16_front(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                             size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, const Front &front) {
    if (front.kind == kFrontSoftF32)
        return launch_i16<RATE, M, WPT, MINB, kFrontSoftF32>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
    return launch_i16<RATE, M, WPT, MINB, kFrontNone>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
}

}  // namespace

// LABRADOR_LDPC_TM_I16_WIDE=1 keeps i16 on the scalar-lane kernel (A/B runs and tests).
bool has_decode_ms_tm_i16(int code) {
    static const bool off = [] { const char *e = getenv("LABRADOR_LDPC_TM_I16_WIDE"); return e && atoi(e) != 0; }();
    return !off && code >= 3 && code <= 8;
}

// Returns true (and launches) if the packed-check-side kernel covers (code, i16, front).
bool launch_decode_ms_tm_i16(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                             uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err, const Front &front) {
    if (!has_decode_ms_tm_i16(code) || front.kind == kFrontHard) return false;
    const CodeInfo &c = *code_info(code);
    switch (code) {
        case 3:      // TM1280: M = 128, 64 threads per codeword, compiled for six resident CTAs per SM like the i8 kernel
            if (!structure_matches<2>(c) || c.m != 128) return false;
            *err = launch_i16_front<2, 128, 1, 6>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
            return true;
        case 4:
            if (!structure_matches<1>(c) || c.m != 256) return false;
            *err = launch_i16_front<1, 256, 1, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
            return true;
        case 5:
            if (!structure_matches<0>(c) || c.m != 512) return false;
            *err = launch_i16_front<0, 512, 1, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
            return true;
        case 6:
            if (!structure_matches<2>(c) || c.m != 512) return false;
            *err = launch_i16_front<2, 512, 1, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
            return true;
        case 7:
            if (!structure_matches<1>(c) || c.m != 1024) return false;
            *err = launch_i16_front<1, 1024, 1, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
            return true;
        case 8:
            if (!structure_matches<0>(c) || c.m != 2048) return false;
            *err = launch_i16_front<0, 2048, 2, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
            return true;
        default:
            return false;
    }
}

}  // namespace ldpc
