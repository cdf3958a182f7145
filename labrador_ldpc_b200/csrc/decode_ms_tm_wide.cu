// Specialised min-sum decoder for the TM codes with one element per lane: i16 / i32 / f32 / f64 LLRs
// (and every LLR type on TM1280, whose M = 128 is too small for the packed-lane i8 kernel).
//
// Replaces LDPCCode::decode_ms::<T> (reference src/decoder.rs:347-475).  Same skeleton as
// decode_ms_tm.cu -- thread t owns element t of every prototype column and every prototype row,
// identity-block messages stay in registers, pi_k-block messages go through shared memory in check
// order, u is never stored, two-stage exit test (row 0 in the threads that own its checks, rows 1-2 from bit-packed
// hard decisions only when row 0 is clean) -- with one value per 32-bit register.
// i8 / i16 use the biased one-instruction arithmetic of biased_arith.cuh; the other types the plain scalar
// DecodeFrom semantics of llr_arith.cuh (reference src/decoder.rs:42-86):
//   variable side  va = llr (+) u_0 (+) u_1 ...  in ascending edge index (:408), then v_j = va (-) u_j (:421)
//   check side     self-correction against the previous v kept in a register (:422-426), min over the
//                  other edges by prefix/suffix minima (= min1/min2 selection, :391-395), sign parity of
//                  the other edges (:398-405).
#include <cuda_runtime.h>

#include <cstdlib>

#include "biased_arith.cuh"
#include "bulk_copy.cuh"
#include "front.cuh"
#include "llr_arith.cuh"
#include "runtime.h"
#include "tm_common.cuh"

namespace ldpc {
using namespace tm;

namespace {

constexpr int kMaxDegW = 18;

// bytes one frame occupies in the caller's buffer
template <class T, int FRONT> __host__ __device__ constexpr unsigned wide_frame_bytes(int n) {
    return FRONT == kFrontSoftF32 ? (unsigned)n * 4u : FRONT == kFrontHard ? (unsigned)n / 8u : (unsigned)n * (unsigned)sizeof(T);
}
// STAGE: the next frame's channel values are fetched into shared memory by a bulk asynchronous copy (TMA) while the
// current frame is decoded, as in decode_ms_tm.cu; chosen by the launcher where two frames fit beside the messages.
template <int RATE, int M, class T, int NT, int FRONT = kFrontNone, bool STAGE = false>
__device__ __forceinline__ void
decode_ms_tm_wide_body(const TmParams &prm, const typename FrontSrc<FRONT, T>::type *__restrict__ llrs_all,
                         uint8_t *__restrict__ out_all,
                         unsigned long long batch, unsigned max_iters, uint8_t *__restrict__ success,
                         uint32_t *__restrict__ iters_out, unsigned long long *__restrict__ counter,
                         const float fscale, const float flimit) {
    typedef Proto<RATE> P;
    typedef Arith<T> A;
    typedef typename MsgStore<T>::type ST;          // shared-memory storage type of a message
    typedef typename A::C CT;                       // register (compute) type
    constexpr int NB = P::NB, NCOL = P::NCOL, NROW = P::NROW;
    constexpr int NP = count_p<P>(NB), NI = NB - NP;
    constexpr int Q = M / 4, EPT = M / NT;
    constexpr int NV = NCOL * M, N = (NCOL - 1) * M, NC = NROW * M;
    constexpr int HBW = NV / 32, SYW = NC / 32;
    constexpr int CA = P::blk(0).col, CP = NCOL - 1;      // the two columns row 0 touches (see decode_ms_tm.cu)
    // i8 / i16: the biased representation of the packed i8 kernel (decode_ms_tm.cu), one value per 32-bit register:
    //   marginal VA = va + B in [0, 2B-1], saturating_add = one VIADDMNMX.RELU;  message C = MAXV - clamp(va - u, +-MAXV);
    //   check side: sign = bit BITS-1 of C, |v| = |C - MAXV|, kill = that bit of (C ^ old) & (C ^ (old + 1)).
    constexpr bool kBiased = std::is_same<T, int8_t>::value || std::is_same<T, int16_t>::value;
    // only where the registers run out: 2 registers per f64 value of per-edge / per-column state against the budget
    constexpr bool kPackOld = std::is_same<T, double>::value &&
                              2 * (NB + (NB - count_p<P>(NB)) + NCOL) * (M / NT) + 40 > (65536 / NT > 255 ? 255 : 65536 / NT);
    constexpr int BITS = 8 * (int)sizeof(T);
    constexpr int kB = kBiased ? (1 << (BITS - 1)) : 0, kMaxV = kB - 1;
    static_assert(M % NT == 0 && NT % 32 == 0 && Q % 32 == 0, "whole warps per quarter");
    static_assert(SYW <= NT && NCOL - 2 <= 16, "one thread per syndrome word; packed hard bits fit");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    ST *msg = reinterpret_cast<ST *>(smem_raw);                                 // [NP][M], check order
    uint32_t *hb = reinterpret_cast<uint32_t *>(smem_raw + sizeof(ST) * NP * M);   // [HBW] packed hard decisions
    // Row 0 of every TM prototype is I(CA) + I(CP) + P(CP) (static_assert'ed in decode_ms_tm.cu): the two identity terms of
    // a check are marginals of the thread that owns the check, the permuted one arrives through this message-shaped byte
    // array (written with the address of block 2), so every thread tests its own row-0 checks inside the check phase.
    uint8_t *hmsg = reinterpret_cast<uint8_t *>(hb + HBW);                         // [M] hard bits of column CP, permuted by block 2
    constexpr int PS2 = count_p<P>(2);
    static_assert(P::blk(2).row == 0 && P::blk(2).col == CP && P::blk(2).isp && !P::blk(0).isp && !P::blk(1).isp &&
                  P::blk(0).col == CA && P::blk(1).col == CP && P::blk(3).row == 1, "row 0 must be I(CA) + I(CP) + P(CP)");
    constexpr unsigned FB = wide_frame_bytes<T, FRONT>(N);
    unsigned char *stage = smem_raw + ((sizeof(ST) * NP * M + sizeof(uint32_t) * HBW + M + 15) & ~(size_t)15);   // [2][FB] if STAGE
    const unsigned char *in_all = reinterpret_cast<const unsigned char *>(llrs_all);
    __shared__ unsigned long long s_frame[2];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x, lane = tid & 31;

    // variable-side address of every pi_k block: check element (q, x) <-> variable element
    // ((theta + q) mod 4, (phi_q + x) mod Q)  (reference src/codes/mod.rs:312-322), inverted here
    int paddr[NP > 0 ? NP : 1][EPT];
#pragma unroll
    for (int ei = 0; ei < EPT; ei++) {
        const int e = tid + ei * NT, qv = e / Q, xv = e % Q;
        static_for<0, NB>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            if constexpr (P::blk(b).isp) {
                constexpr int ps = count_p<P>(b);
                const int q = (qv - (int)prm.theta[b]) & 3;
                const int x = (xv - (int)prm.phi[b][q]) & (Q - 1);
                paddr[ps][ei] = ps * M + q * Q + x;
            }
        });
    }

    // Frames are claimed one ahead; with STAGE the copy of the next frame is in flight while this one is decoded.
    const bool use_bulk = STAGE && FB % 16 == 0 && (reinterpret_cast<uintptr_t>(llrs_all) & 15u) == 0;
    if (tid == 0) {
        if constexpr (STAGE) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            mbar_init_fence();
        }
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
        const typename FrontSrc<FRONT, T>::type *llr;
        if (use_bulk) {
            mbar_wait(&s_bar[cur], (bar_parity >> cur) & 1u);
            bar_parity ^= 1u << cur;
            llr = reinterpret_cast<const typename FrontSrc<FRONT, T>::type *>(stage + cur * FB);
        } else {
            llr = reinterpret_cast<const typename FrontSrc<FRONT, T>::type *>(in_all + frame * (unsigned long long)FB);
        }

        // zero-initialised state, every call (:368, :374)
        CT Lv[NCOL][EPT], idm[NI > 0 ? NI : 1][EPT], vold[kPackOld ? 1 : NB][EPT];
        // f64 on the k = 4096 codes: the self-correction rule (:422-426) needs only "old v negative" and "old v non-zero"
        // of every edge; two bits instead of a 64-bit register per edge keep them (almost) out of local memory
        uint32_t oldneg[(NB + 31) / 32][EPT], oldnz[(NB + 31) / 32][EPT];
#pragma unroll
        for (int ei = 0; ei < EPT; ei++) {
            const int e = tid + ei * NT;
#pragma unroll
            for (int c = 0; c < NCOL; c++) {
                Lv[c][ei] = c < NCOL - 1 ? front_load<FRONT, T>(llr, c * M + e, fscale, flimit) : A::zero();   // :382-383
                if constexpr (kBiased) Lv[c][ei] += (CT)kB;
            }
#pragma unroll
            for (int i = 0; i < NI; i++) idm[i][ei] = A::zero();
#pragma unroll
            for (int b = 0; b < (kPackOld ? 1 : NB); b++) vold[b][ei] = kBiased ? (CT)kMaxV : A::zero();
#pragma unroll
            for (int w = 0; w < (NB + 31) / 32; w++) { oldneg[w][ei] = 0; oldnz[w][ei] = 0; }
#pragma unroll
            for (int p = 0; p < NP; p++) msg[p * M + e] = (ST)A::zero();
        }
        for (int i = tid; i < HBW; i += NT) hb[i] = 0;
        __syncthreads();

        unsigned iters_run = max_iters;
        bool ok = false, hb_complete = true;
        uint32_t pack[EPT];            // hard bits of every column of the latest variable phase (column NCOL-1 at bit 0)
        bool hloc[EPT], bad[EPT];      // XOR of the two in-thread terms of the thread's row-0 check; its parity
        auto flush_pack = [&]() {      // ballot-pack the hard bits of all columns (second stage / output)
#pragma unroll
            for (int ei = 0; ei < EPT; ei++) {
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = NCOL - 1 - decltype(ci)::value;
                    const unsigned bw = __ballot_sync(0xFFFFFFFFu, (pack[ei] >> decltype(ci)::value) & 1u);
                    if (lane == 0) hb[(c * M + tid + ei * NT) >> 5] = bw;
                });
            }
        };

        for (unsigned iter = 0; iter < max_iters; iter++) {
            // ================= variable phase =================
#pragma unroll
            for (int ei = 0; ei < EPT; ei++) {
                pack[ei] = 0;
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    CT va = Lv[c][ei];
                    CT ub[6];
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            CT u;
                            if constexpr (P::blk(b).isp) u = (CT)msg[paddr[count_p<P>(b)][ei]];
                            else u = idm[count_i<P>(b)][ei];
                            ub[k] = u;
                            if constexpr (kBiased) va = (CT)__viaddmin_s32_relu((int)va, (int)u, 2 * kB - 1);
                            else va = A::sat_add(va, u);                              // :408, ascending idx
                        }
                    });
                    bool hard;
                    if constexpr (kBiased) hard = (int)va < kB;
                    else hard = A::hard_bit(va);
                    [[maybe_unused]] const int van = kBiased ? (2 * kB - 1) - (int)va : 0;
                    pack[ei] = pack[ei] * 2u + (hard ? 1u : 0u);
                    if constexpr (c == CA) hloc[ei] = hard;
                    if constexpr (c == CP) {
                        hloc[ei] ^= hard;
                        hmsg[paddr[PS2][ei] - PS2 * M] = hard ? 1 : 0;
                    }
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            CT nv;
                            if constexpr (kBiased) nv = (CT)__viaddmin_s32_relu(van, (int)ub[k], 2 * kMaxV);
                            else nv = A::sat_sub(va, ub[k]);                          // :421
                            if constexpr (P::blk(b).isp) msg[paddr[count_p<P>(b)][ei]] = (ST)nv;
                            else idm[count_i<P>(b)][ei] = nv;
                        }
                    });
                });
            }
            __syncthreads();

            // ================= check phase =================
#pragma unroll
            for (int ei = 0; ei < EPT; ei++) {
                const int e = tid + ei * NT;
                bad[ei] = hloc[ei] != (hmsg[e] != 0);            // parity of the three marginals of row-0 check e (:445-447)
                static_for<0, NROW>([&](auto ri) {
                    constexpr int r = decltype(ri)::value;
                    constexpr int DC = row_degree<P>(r);
                    if constexpr (kBiased) {
                        uint32_t a[kMaxDegW], ck[kMaxDegW], mu[kMaxDegW];
                        uint32_t sx = 0;
                        static_for<0, NB>([&](auto bi) {
                            constexpr int b = decltype(bi)::value;
                            if constexpr (P::blk(b).row == r) {
                                constexpr int k = pos_in_row<P>(b);
                                uint32_t cv;
                                if constexpr (P::blk(b).isp) cv = (uint32_t)msg[count_p<P>(b) * M + e];
                                else cv = (uint32_t)idm[count_i<P>(b)][ei];
                                const uint32_t old = (uint32_t)vold[b][ei];
                                const uint32_t x = (cv ^ old) & (cv ^ (old + 1u));       // bit BITS-1: sign flipped and old != 0
                                const uint32_t km = sign_mask_of_byte<BITS / 8 - 1>(x);
                                const uint32_t cor = (cv & ~km) | ((uint32_t)kMaxV & km);   // killed -> v = 0  (:422-426)
                                vold[b][ei] = (CT)cor;
                                ck[k] = cor;
                                a[k] = __usad(cor, (uint32_t)kMaxV, 0u);                 // |v|
                                sx ^= cor;                                               // bit BITS-1: product of signs
                            }
                        });
                        min_excluding_self_u32<DC, kMaxDegW>(a, mu);                               // :391-395
                        static_for<0, NB>([&](auto bi) {
                            constexpr int b = decltype(bi)::value;
                            if constexpr (P::blk(b).row == r) {
                                constexpr int k = pos_in_row<P>(b);
                                const uint32_t nm = sign_mask_of_byte<BITS / 8 - 1>(sx ^ ck[k]);   // u negative  (:398-405)
                                const uint32_t u = (mu[k] + nm) ^ nm;                    // +-mu, two's complement
                                if constexpr (P::blk(b).isp) msg[count_p<P>(b) * M + e] = (ST)u;
                                else idm[count_i<P>(b)][ei] = (CT)u;
                            }
                        });
                    } else {
                    CT a[kMaxDegW], suf[kMaxDegW];
                    bool sg[kMaxDegW];
                    bool stot = false;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            CT nv;
                            if constexpr (P::blk(b).isp) nv = (CT)msg[count_p<P>(b) * M + e];
                            else nv = idm[count_i<P>(b)][ei];
                            CT v;
                            if constexpr (kPackOld) {
                                constexpr int w = b / 32;
                                constexpr uint32_t bit = 1u << (b % 32);
                                const bool keep = (A::hard_bit(nv) == ((oldneg[w][ei] & bit) != 0)) || (oldnz[w][ei] & bit) == 0;
                                v = keep ? nv : A::zero();                            // :422-426
                                oldneg[w][ei] = A::hard_bit(v) ? (oldneg[w][ei] | bit) : (oldneg[w][ei] & ~bit);
                                oldnz[w][ei] = (v == A::zero()) ? (oldnz[w][ei] & ~bit) : (oldnz[w][ei] | bit);
                            } else {
                                const CT vo = vold[b][ei];
                                const bool keep = (A::hard_bit(nv) == A::hard_bit(vo)) || (vo == A::zero());
                                v = keep ? nv : A::zero();                            // :422-426
                                vold[b][ei] = v;
                            }
                            a[k] = A::abs(v);
                            sg[k] = A::hard_bit(v);
                            stot ^= sg[k];                                            // :439-441
                        }
                    });
                    // min over the other edges = (|v| == min1 ? min2 : min1)  (:391-395)
                    // f32 / i32 have a three-input minimum (FMNMX3 / VIMNMX3 on sm_100a, which nested two-input minima
                    // compile to): prefixes and suffixes at pair boundaries, 2 DC - 2 minima instead of 3 DC - 6; f64 has
                    // no minimum instruction at all and keeps the form with the fewest two-input minima
                    constexpr bool kMin3 = std::is_same<T, float>::value || std::is_same<T, int32_t>::value;
                    CT muv[kMaxDegW];
                    if constexpr (kMin3 && DC >= 4) {
                        constexpr int NPAIR = DC / 2;
                        constexpr bool ODD = (DC & 1) != 0;
                        if constexpr (ODD) suf[NPAIR] = a[DC - 1];
#pragma unroll
                        for (int j = NPAIR - 1; j >= 1; j--) {
                            if (j == NPAIR - 1 && !ODD) suf[j] = A::min(a[2 * j], a[2 * j + 1]);
                            else suf[j] = A::min(A::min(a[2 * j], a[2 * j + 1]), suf[j + 1]);
                        }
                        CT pre2 = A::zero();
#pragma unroll
                        for (int j = 0; j < NPAIR; j++) {
                            const bool has_pre = j > 0, has_suf = (j + 1 < NPAIR) || ODD;
                            if (has_pre && has_suf) {
                                muv[2 * j] = A::min(A::min(pre2, a[2 * j + 1]), suf[j + 1]);
                                muv[2 * j + 1] = A::min(A::min(pre2, a[2 * j]), suf[j + 1]);
                            } else if (has_suf) {
                                muv[2 * j] = A::min(a[2 * j + 1], suf[j + 1]);
                                muv[2 * j + 1] = A::min(a[2 * j], suf[j + 1]);
                            } else {
                                muv[2 * j] = A::min(pre2, a[2 * j + 1]);
                                muv[2 * j + 1] = A::min(pre2, a[2 * j]);
                            }
                            if (j + 1 < NPAIR || ODD)
                                pre2 = has_pre ? A::min(A::min(pre2, a[2 * j]), a[2 * j + 1]) : A::min(a[2 * j], a[2 * j + 1]);
                        }
                        if constexpr (ODD) muv[DC - 1] = pre2;
                    } else {
                        suf[DC - 1] = a[DC - 1];
#pragma unroll
                        for (int k = DC - 2; k >= 1; k--) suf[k] = A::min(a[k], suf[k + 1]);
                        CT pre = a[0];
#pragma unroll
                        for (int k = 0; k < DC; k++) {
                            if (k == 0) muv[k] = suf[1];
                            else if (k == DC - 1) muv[k] = pre;
                            else muv[k] = A::min(pre, suf[k + 1]);
                            if (k > 0 && k < DC - 1) pre = A::min(pre, a[k]);
                        }
                    }
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            CT u = muv[k];
                            if (stot != sg[k]) u = A::neg(u);                          // :398-405
                            if constexpr (P::blk(b).isp) msg[count_p<P>(b) * M + e] = (ST)u;
                            else idm[count_i<P>(b)][ei] = u;
                        }
                    });
                    }
                });
            }
            // ---- parity of the marginals' hard bits (:445-453): two-stage, bit-packed ----
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
            for (int ei = 0; ei < EPT; ei++) synd |= bad[ei] ? 1u : 0u;
            hb_complete = false;
            if (__syncthreads_or(synd != 0) == 0) {
                flush_pack();
                hb_complete = true;
                __syncthreads();
                synd = 0;
                for (int sw = M / 32 + tid; sw < SYW; sw += NT) synd |= syndrome_word(sw);
                if (__syncthreads_or(synd != 0) == 0) {
                    ok = true;
                    iters_run = iter;                                                 // :462
                    break;
                }
            }
        }
        if (!hb_complete) {
            flush_pack();
            __syncthreads();
        }

        // output: hard decisions of all n+p marginals, MSB first (:455-461, :466-473)
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

// shared memory of one CTA without / with the two staging buffers
template <int RATE, int M, class T> __host__ __device__ constexpr size_t wide_base_smem() {
    typedef Proto<RATE> P;
    return ((sizeof(typename MsgStore<T>::type) * count_p<P>(P::NB) * M + sizeof(uint32_t) * P::NCOL * M / 32 + M + 15) & ~(size_t)15);
}
template <int RATE, int M, class T, int FRONT> __host__ __device__ constexpr bool wide_stage() {
    constexpr size_t fb = wide_frame_bytes<T, FRONT>((Proto<RATE>::NCOL - 1) * M);
    // Measured on every code (profiles/raw/r02j_wide_staging_log.txt against r02h_wide_kernel_log.txt): +1.5 % on TM2048
    // (f32, i32: the C2 configuration), -1 ... -5.5 % elsewhere (the kernels are ALU-pipe bound, the staging buffers
    // cost residency or L1), so it is enabled where it pays.  LDPC_WIDE_STAGE_ALL (compile time) stages wherever it fits.
#ifdef LDPC_WIDE_STAGE_ALL
    constexpr bool wanted = true;
#else
    constexpr bool wanted = RATE == 0 && M == 512;
#endif
    return wanted && fb % 16 == 0 && wide_base_smem<RATE, M, T>() + 2 * fb <= 200 * 1024;
}

// Two entry points over the same body.  __launch_bounds__(NT) lets ptxas trade registers for occupancy, which is right
// for the 32-bit types (TM2048 f32: 64 registers, two CTAs per SM, 10.3 M cw/s; with the full register budget 8.4 M).
// For f64 it picked 32-64 registers and spilled 1.5-2 KB per thread on the k = 4096 codes; __maxnreg__ hands it the
// whole budget of one CTA per SM instead (TM5120 f64 0.46 -> 2.3 M cw/s, TM6144 0.53 -> 1.5 M).
template <int RATE, int M, class T, int NT, int FRONT = kFrontNone>
__global__ void __launch_bounds__(NT)
decode_ms_tm_wide_kernel(const TmParams prm, const typename FrontSrc<FRONT, T>::type *__restrict__ llrs_all,
                         uint8_t *__restrict__ out_all, unsigned long long batch, unsigned max_iters,
                         uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out,
                         unsigned long long *__restrict__ counter, const float fscale, const float flimit) {
    decode_ms_tm_wide_body<RATE, M, T, NT, FRONT, wide_stage<RATE, M, T, FRONT>()>(prm, llrs_all, out_all, batch, max_iters, success,
                                                                                    iters_out, counter, fscale, flimit);
}
template <int RATE, int M, class T, int NT, int FRONT = kFrontNone>
__global__ void __maxnreg__(65536 / NT > 255 ? 255 : 65536 / NT)
decode_ms_tm_wide_kernel_allregs(const TmParams prm, const typename FrontSrc<FRONT, T>::type *__restrict__ llrs_all,
                                 uint8_t *__restrict__ out_all, unsigned long long batch, unsigned max_iters,
                                 uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out,
                                 unsigned long long *__restrict__ counter, const float fscale, const float flimit) {
    decode_ms_tm_wide_body<RATE, M, T, NT, FRONT, wide_stage<RATE, M, T, FRONT>()>(prm, llrs_all, out_all, batch, max_iters, success,
                                                                                    iters_out, counter, fscale, flimit);
}

template <int RATE, int M, class T, int NT, int FRONT = kFrontNone>
cudaError_t launch_wide(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                        size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                        const Front &front = Front()) {
    typedef Proto<RATE> P;
    constexpr int NP = count_p<P>(P::NB);
    const TmParams prm = make_params<RATE>(c);
    // messages, hard-bit words, hmsg (+ two frames of staging where they fit)
    const size_t smem = wide_base_smem<RATE, M, T>() +
                        (wide_stage<RATE, M, T, FRONT>() ? 2 * (size_t)wide_frame_bytes<T, FRONT>((P::NCOL - 1) * M) : 0);
    auto kern = [] {
        if constexpr (std::is_same<T, double>::value) return &decode_ms_tm_wide_kernel_allregs<RATE, M, T, NT, FRONT>;
        else return &decode_ms_tm_wide_kernel<RATE, M, T, NT, FRONT>;
    }();
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
    unsigned long long *counter = wc.ptr();
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)grid, NT, smem, stream>>>(prm, static_cast<const typename FrontSrc<FRONT, T>::type *>(llrs), output,
                                               (unsigned long long)batch, mi, success, iters, counter, front.scale,
                                               front.limit);
    count_launch();
    return cudaGetLastError();
}

template <int RATE, int M, int NT, bool WITH_I8>
bool dispatch_type(DeviceCtx &ctx, const CodeInfo &c, int llr_type, const void *llrs, uint8_t *output,
                   size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                   cudaError_t *err, const Front &front) {
    if (front.kind != kFrontNone) {      // fused front ends: (soft, i8 | i16) and (hard, i8)
        if (front.kind == kFrontSoftF32 && llr_type == kI16) {
            *err = launch_wide<RATE, M, int16_t, NT, kFrontSoftF32>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
            return true;
        }
        if constexpr (WITH_I8) {
            if (front.kind == kFrontSoftF32 && llr_type == kI8) {
                *err = launch_wide<RATE, M, int8_t, NT, kFrontSoftF32>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
                return true;
            }
            if (front.kind == kFrontHard && llr_type == kI8) {
                *err = launch_wide<RATE, M, int8_t, NT, kFrontHard>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
                return true;
            }
        }
        return false;
    }
    switch (llr_type) {
        case kI8:
            if constexpr (WITH_I8) {
                *err = launch_wide<RATE, M, int8_t, NT>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
                return true;
            } else {
                return false;
            }
        case kI16:
            *err = launch_wide<RATE, M, int16_t, NT>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
            return true;
        case kI32:
            *err = launch_wide<RATE, M, int32_t, NT>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
            return true;
        case kF32:
            *err = launch_wide<RATE, M, float, NT>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
            return true;
        case kF64:
            *err = launch_wide<RATE, M, double, NT>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
            return true;
        default:
            return false;
    }
}

}  // namespace

// Returns true (and launches) if the wide-lane TM kernel covers (code, llr_type).
bool launch_decode_ms_tm_wide(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output, size_t batch,
                              size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                              cudaError_t *err, const Front &front) {
    const CodeInfo &c = *code_info(code);
    switch (code) {
        case 3:   // TM1280: every type (M = 128 is too small for the packed i8 kernel)
            if (!structure_matches<2>(c) || c.m != 128) return false;
            return dispatch_type<2, 128, 128, true>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 4:
            if (!structure_matches<1>(c) || c.m != 256) return false;
            return dispatch_type<1, 256, 256, false>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 5:
            if (!structure_matches<0>(c) || c.m != 512) return false;
            return dispatch_type<0, 512, 512, false>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 6:
            if (!structure_matches<2>(c) || c.m != 512) return false;
            return dispatch_type<2, 512, 512, false>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 7:
            if (!structure_matches<1>(c) || c.m != 1024) return false;
            return dispatch_type<1, 1024, 512, false>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 8:
            if (!structure_matches<0>(c) || c.m != 2048) return false;
            if (llr_type == kF64 && front.kind == kFrontNone) {      // 512 threads x 128 registers instead of 1024 x 64
                *err = launch_wide<0, 2048, double, 512>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
                return true;
            }
            return dispatch_type<0, 2048, 1024, false>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        default:
            return false;
    }
}

bool has_decode_ms_tm_wide(int code, int llr_type) {
    if (code < 3 || code > 8) return false;
    if (llr_type == kI8) return code == 3;
    return true;
}

}  // namespace ldpc
