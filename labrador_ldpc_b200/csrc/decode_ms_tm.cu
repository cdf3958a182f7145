// Specialised min-sum decoder for the TM codes with i8 LLRs (the headline path).
//
// Replaces LDPCCode::decode_ms::<i8> (reference src/decoder.rs:347-475) for
// TM1536/TM2048/TM5120/TM6144/TM8192; results (decoded bytes, success flag,
// iteration count) are bit-identical to the reference's flooding schedule.
//
// One CTA decodes one codeword at a time (persistent CTAs pull frame indices
// from an atomic counter).  All arithmetic is done on two 16-bit lanes per
// 32-bit register with the native packed instructions of sm_100a
// (VIADDMNMX.S16x2.RELU, VIMNMX.S16x2, VIADD.16x2, PRMT, LOP3).
//
// Lane pairing.  Inside every quarter (Q = M/4 elements) of an MxM block the
// elements x and x + S (S = Q/2) share a register.  The pi_k permutation of a
// block is "quarter q -> (theta+q) mod 4, offset x -> (phi_q + x) mod Q"
// (reference src/codes/mod.rs:312-322), so it maps lane pairs to lane pairs --
// at most the two lanes swap.  Word slot = q*S + w, w in [0,S).  Thread t owns
// word slot t of EVERY prototype column (variable side) and EVERY prototype
// row (check side); M/2 threads per CTA.
//
// Messages.  Identity blocks connect check slot t with variable slot t of the
// same thread, so their messages never leave the thread's registers.  Only the
// permutation blocks go through shared memory: one 32-bit word per lane pair,
// stored in check order (own slot for the check side, permuted address +
// optional lane swap for the variable side; consecutive threads hit
// consecutive words, so there are no bank conflicts).
//
// Representations (all exact, see DESIGN.md for the derivations):
//   variable side: VA = marginal + 128 in [0,255]; saturating_add(va, u) is one
//       VIADDMNMX.RELU: max(min(VA + u, 255), 0)                     (:408)
//   variable -> check: C = 127 - clamp(va - u, -127, 127) in [0,254], one
//       VIADDMNMX.RELU on VAn = 255 - VA.  -128 and -127 are interchangeable
//       because v is only ever used through saturating_abs, sign and == 0.
//   check side: sign = bit 7 of C, |v| + 127 = max(C, 254 - C).  The
//       self-correction rule (:422-426) needs only the sign class of the
//       previous v, kept as Cc_old: kill = bit7((C ^ Cc_old) & (C ^ (Cc_old+1))).
//   check -> variable: u_j = +-(min_{k != j} |v_k|) by prefix/suffix minima
//       (equal to the reference's min1/min2 selection, :391-395), sign
//       = parity of the other edges' signs (:398-405), as two's complement.
// The per-iteration parity test (:445-453) runs bit-packed: warp ballots pack
// the marginals' hard bits, the syndrome is XORs of funnel-shifted words.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "bulk_copy.cuh"
#include "front.cuh"
#include "runtime.h"
#include "tm_common.cuh"

namespace ldpc {

using namespace tm;

namespace {

// 0xFFFF in every 16-bit lane whose bit 15 is set
__device__ __forceinline__ uint32_t prmt_sign15(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xbb99;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t prmt_sign7(uint32_t x) {
    // bytes 0,1 <- sign of byte 0; bytes 2,3 <- sign of byte 2 (bit 7 of each 16-bit lane -> lane mask)
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xaa88;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t lane_rot(uint32_t x, uint32_t sh) { return __funnelshift_l(x, x, sh); }

constexpr int kMaxDeg = 18;

// ---- FMA-pipe arithmetic on small integers held as fp16 bit patterns ----
// A 16-bit lane holding the integer k (0 <= k < 1024) IS the fp16 subnormal k * 2^-24, and a set bit 15
// makes it -k.  fp16 add / fma on such values is exact integer arithmetic (|result| < 2048), runs on the
// FMA pipe (HFMA2 / HADD2) and leaves the saturated ALU pipe alone (profiles/r01_pipe_ubench.md).
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint32_t f_relu_add(uint32_t a, uint32_t b) {          // max(a + b, 0)
    return h2u(__hfma2_relu(u2h(a), u2h(0x3C003C00u), u2h(b)));
}
__device__ __forceinline__ uint32_t f_relu_sub(uint32_t a, uint32_t b) {          // max(a - b, 0)
    return h2u(__hfma2_relu(u2h(b), u2h(0xBC00BC00u), u2h(a)));
}
__device__ __forceinline__ uint32_t f_sub(uint32_t a, uint32_t b) { return h2u(__hsub2(u2h(a), u2h(b))); }
// min of two non-negative lanes: on the FMA pipe (2 ops: a - relu(a - b)) or the ALU pipe (1 VIMNMX)
template <bool ON_FMA> __device__ __forceinline__ uint32_t pmin(uint32_t a, uint32_t b) {
    if constexpr (ON_FMA) return f_sub(a, f_relu_sub(a, b));
    else return __vminu2(a, b);
}

// Minimum over "all the other edges" for every edge of one check word: prefix/suffix minima.
// SUF_F / PRE_F / COMB_F choose the pipe of the three groups of DC-2 minima (pipe balancing).
// All-integer form with the three-input minimum (VIMNMX3.U16x2): prefixes P_j = min(a_0..a_{2j-1}) and suffixes
// S_j = min(a_{2j}..) at pair boundaries only, then mu_{2j} = min3(P_j, a_{2j+1}, S_{j+1}) and
// mu_{2j+1} = min3(P_j, a_{2j}, S_{j+1}):  2*DC - 2 instructions instead of 3*DC - 6.
template <int DC>
__device__ __forceinline__ void min_excluding_self3(const uint32_t (&a)[kMaxDeg], uint32_t (&mu)[kMaxDeg]) {
    constexpr int NPAIR = DC / 2;                  // full pairs; an odd DC leaves a[DC-1] on its own
    constexpr bool ODD = (DC & 1) != 0;
    uint32_t suf[kMaxDeg / 2 + 2];                 // suf[j] = min(a[2j] .. a[DC-1]), j = 1 .. NPAIR-1 (+ the odd tail)
    if constexpr (ODD) suf[NPAIR] = a[DC - 1];
#pragma unroll
    for (int j = NPAIR - 1; j >= 1; j--) {
        if (j == NPAIR - 1 && !ODD) suf[j] = __vminu2(a[2 * j], a[2 * j + 1]);
        else suf[j] = __vimin3_u16x2(a[2 * j], a[2 * j + 1], suf[j + 1]);
    }
    uint32_t pre = 0;                              // pre = min(a[0] .. a[2j-1]), defined from j = 1
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
        } else {                                   // DC == 2
            mu[2 * j] = a[2 * j + 1];
            mu[2 * j + 1] = a[2 * j];
        }
        if (j + 1 < NPAIR || ODD)
            pre = has_pre ? __vimin3_u16x2(pre, a[2 * j], a[2 * j + 1]) : __vminu2(a[2 * j], a[2 * j + 1]);
    }
    if constexpr (ODD) mu[DC - 1] = pre;
}

// The same on fp16 lanes: a lane holding the signed integer d (|d| < 1024) as a sign-magnitude subnormal (bit 15 = sign,
// low bits = |d|).  HMNMX2 / VHMNMX take |.| as a free operand modifier, so no separate absolute value is needed, and the
// result (a non-negative subnormal) is bit-identical to the integer |d|.  TWO_ONLY: 2-input minima only (3*DC-6 full-rate
// HMNMX2) instead of the 2*DC-2 mix with the three-input VHMNMX.
__device__ __forceinline__ uint32_t hmin2a(uint32_t x, uint32_t y) { return h2u(__hmin2(__habs2(u2h(x)), __habs2(u2h(y)))); }
__device__ __forceinline__ uint32_t hmin3a(uint32_t x, uint32_t y, uint32_t z) {
    return h2u(__hmin2(__hmin2(__habs2(u2h(x)), __habs2(u2h(y))), __habs2(u2h(z))));
}
template <int DC, bool TWO_ONLY>
__device__ __forceinline__ void min_excluding_self_h(const uint32_t (&a)[kMaxDeg], uint32_t (&mu)[kMaxDeg]) {
    static_assert(DC >= 3, "every output must come out of a minimum (which applies the absolute value)");
    if constexpr (TWO_ONLY) {
        uint32_t suf[kMaxDeg];
        suf[DC - 1] = a[DC - 1];
#pragma unroll
        for (int k = DC - 2; k >= 1; k--) suf[k] = hmin2a(a[k], suf[k + 1]);
        uint32_t pre = a[0];
        mu[0] = suf[1];
#pragma unroll
        for (int k = 1; k < DC - 1; k++) {
            mu[k] = hmin2a(pre, suf[k + 1]);
            pre = hmin2a(pre, a[k]);
        }
        mu[DC - 1] = pre;
    } else {
        constexpr int NPAIR = DC / 2;
        constexpr bool ODD = (DC & 1) != 0;
        uint32_t suf[kMaxDeg / 2 + 2];
        if constexpr (ODD) suf[NPAIR] = a[DC - 1];
#pragma unroll
        for (int j = NPAIR - 1; j >= 1; j--) {
            if (j == NPAIR - 1 && !ODD) suf[j] = hmin2a(a[2 * j], a[2 * j + 1]);
            else suf[j] = hmin3a(a[2 * j], a[2 * j + 1], suf[j + 1]);
        }
        uint32_t pre = 0;
#pragma unroll
        for (int j = 0; j < NPAIR; j++) {
            const bool has_pre = j > 0, has_suf = (j + 1 < NPAIR) || ODD;
            if (has_pre && has_suf) {
                mu[2 * j] = hmin3a(pre, a[2 * j + 1], suf[j + 1]);
                mu[2 * j + 1] = hmin3a(pre, a[2 * j], suf[j + 1]);
            } else if (has_suf) {
                mu[2 * j] = hmin2a(a[2 * j + 1], suf[j + 1]);
                mu[2 * j + 1] = hmin2a(a[2 * j], suf[j + 1]);
            } else {
                mu[2 * j] = hmin2a(pre, a[2 * j + 1]);
                mu[2 * j + 1] = hmin2a(pre, a[2 * j]);
            }
            if (j + 1 < NPAIR || ODD)
                pre = has_pre ? hmin3a(pre, a[2 * j], a[2 * j + 1]) : hmin2a(a[2 * j], a[2 * j + 1]);
        }
        if constexpr (ODD) mu[DC - 1] = pre;
    }
}

template <int DC, bool SUF_F, bool PRE_F, bool COMB_F>
__device__ __forceinline__ void min_excluding_self(const uint32_t (&a)[kMaxDeg], uint32_t (&mu)[kMaxDeg]) {
    uint32_t suf[kMaxDeg];
    suf[DC - 1] = a[DC - 1];
#pragma unroll
    for (int k = DC - 2; k >= 1; k--) suf[k] = pmin<SUF_F>(a[k], suf[k + 1]);
    uint32_t pre = a[0];
    mu[0] = suf[1];
#pragma unroll
    for (int k = 1; k < DC - 1; k++) {
        mu[k] = pmin<COMB_F>(pre, suf[k + 1]);
        pre = pmin<PRE_F>(pre, a[k]);
    }
    mu[DC - 1] = pre;
}

// Arithmetic variants of the kernel (kept side by side for A/B measurement):
//   ARITH 1: integer lanes throughout (VIADDMNMX chain, two's-complement u)      -- ALU-pipe bound
//   ARITH 3: ARITH 1 with |v| by VABSDIFF4 (no +127 offset to carry around)
//   ARITH 5: ARITH 3 with the three-input minimum (2*DC-2 instead of 3*DC-6 minima per check word)
//   ARITH 6: ARITH 5 with |v| and the minima on fp16 lanes: d = C - 127 by one HADD2 (FMA pipe) instead of VABSDIFF4,
//            |d| as the free operand modifier of HMNMX2 / VHMNMX (KNOBS bit 4: two-input minima only)
//   ARITH 7: ARITH 6 with the sign-magnitude u of ARITH 4
//   ARITH 8: ARITH 6 with the self-correction rule on fp16 lanes too (FMA pipe).  The variable side sends 0x6400 + C
//            (as fp16: 1024 + C, one IMAD); v = 1151 - (1024 + C) by one HADD2 is an integer-valued fp16;
//            keep = sat(v v_old + 1) (one saturating HFMA2) is 0 exactly where the sign flipped and v_old != 0;
//            v_cor = v keep + 0 (one HFMA2); mu * 2^-24 (one HMUL2) is the integer the variable side adds.
//            Five FMA-pipe instructions replace IMAD + LOP3 + PRMT + LOP3 + HADD2; signs are bit 15 of the lanes
//   ARITH 9: ARITH 8 with the sign of u from an fp16 product (HMUL2 of a +-0 word and v_cor) instead of a LOP3
//   ARITH 11: ARITH 10 with the lane swap of a permuted message done by a PRMT with a per-thread selector instead of a
//            funnel shift with a per-thread amount: on the way in the same PRMT also writes the 0x64 exponent bytes
//            (one instruction instead of IMAD + SHF)
//   ARITH 10: ARITH 8 with u = +-mu by one HFMA2: mu * (+-1.0) + 1536 has the bit pattern 0x6600 +- mu, a per-lane
//            subtraction leaves the two's-complement message (LOP3 + HFMA2 + VIADD.16x2 instead of
//            HMUL2 + PRMT + HMUL2 + VIADD.16x2 + LOP3)
//   ARITH 4: ARITH 3 with u sent in sign-magnitude (1 LOP3 + 1 IMAD on the check side) and converted to
//            two's complement on the FMA pipe by the variable side (fp16 magic-constant add); KNOBS as ARITH 2
//   ARITH 2: variable side in fp16 on the FMA pipe, u in sign-magnitude, |v| by VABSDIFF4,
//            part of the minima on the FMA pipe (bits of KNOBS: 1 cv-min, 2 suffix, 4 prefix, 8 combine)
// FRONT (front.cuh): what a frame of `llrs_all` holds -- n int8 LLRs, n float soft values quantised on load,
// or n/8 bytes of hard decisions; the frame is staged in shared memory by the same bulk copy in every case.
// PROF (development aid, tools/tm_phase_prof.py): every warp accumulates the clock cycles it spends in the variable
// phase, at the barrier after it, in the check phase and in the exit test (incl. its barriers) into prof[warp][4].
template <int RATE, int M, int WPT, int ARITH, int KNOBS, int MINB = 1, int FRONT = kFrontNone, bool PROF = false>
__global__ void __launch_bounds__(M / 2 / WPT, MINB)
decode_ms_tm_i8_kernel(const TmParams prm, const typename FrontSrc<FRONT, int8_t>::type *__restrict__ llrs_all,
                       uint8_t *__restrict__ out_all,
                       unsigned long long batch, unsigned max_iters, uint8_t *__restrict__ success,
                       uint32_t *__restrict__ iters_out, unsigned long long *__restrict__ counter,
                       const uint32_t one /* == 1: keeps the borrow-free subtractions on the FMA pipe (IMAD) */,
                       const float fscale, const float flimit, unsigned long long *__restrict__ prof) {
    typedef typename FrontSrc<FRONT, int8_t>::type Src;
    typedef Proto<RATE> P;
    constexpr int NB = P::NB, NCOL = P::NCOL, NROW = P::NROW;
    constexpr int NP = count_p<P>(NB), NI = NB - NP;
    constexpr int Q = M / 4, S = Q / 2, NT = M / 2 / WPT;
    constexpr int NV = NCOL * M, N = (NCOL - 1) * M, NC = NROW * M;
    constexpr int HBW = NV / 32;           // hard-decision words
    constexpr int SYW = NC / 32;           // syndrome words
    // the ballot-packed hard bits (exit test without KNOBS bit 5) need a whole warp per half quarter; with the in-thread
    // exit test only the rare second stage packs bits, and does it with atomics when S < 32 (TM1280: M = 128, S = 16)
    static_assert((KNOBS & 32) != 0 ? (S >= 16 && NT % 32 == 0) : (S >= 32 && S % 32 == 0),
                  "lane pairing needs at least one warp per half quarter");
    static_assert(SYW <= NT, "one thread per syndrome word");
    // Two-stage exit test.  Row 0 of every TM prototype is {I at column CA, I and P at the punctured column CP}.
    // Stage 1 (every iteration) packs the hard bits of those two columns only and tests the M row-0 checks; a
    // non-zero row-0 syndrome proves the codeword is not yet valid.  Only when row 0 is clean are the remaining
    // columns' hard bits (kept as one packed register per thread) ballot-packed and rows 1..2 tested.
    constexpr int CA = P::blk(0).col, CP = NCOL - 1;
    static_assert(P::blk(0).row == 0 && !P::blk(0).isp && P::blk(1).row == 0 && P::blk(1).col == CP && !P::blk(1).isp &&
                  P::blk(2).row == 0 && P::blk(2).col == CP && P::blk(2).isp && P::blk(3).row == 1,
                  "row 0 must be I(CA) + I(CP) + P(CP)");
    static_assert(NCOL - 2 <= 9, "packed hard bits use bits 7..15 / 23..31");

    extern __shared__ __align__(16) uint32_t smem_u32[];
    uint32_t *msg = smem_u32;                       // [NP][M/2] permutation-block messages, check order
    uint32_t *hb = msg + NP * (M / 2);              // [HBW] packed hard decisions, bit i of word j = variable 32j+i
    constexpr unsigned FB = FRONT == kFrontSoftF32 ? N * 4 : FRONT == kFrontHard ? N / 8 : N;   // input bytes per frame
    static_assert(FB % 16 == 0, "bulk copies move multiples of 16 bytes");
    unsigned char *stage = reinterpret_cast<unsigned char *>(hb + ((HBW + 3) & ~3));   // [2][FB] input staging (bulk-copy destination)
    const unsigned char *in_all = reinterpret_cast<const unsigned char *>(llrs_all);
    __shared__ unsigned long long s_frame[2];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31;

    // per-thread constants: permuted word address + lane swap of every permutation block
    uint32_t paddr[NP > 0 ? NP : 1][WPT], pswp[NP > 0 ? NP : 1][WPT];
#pragma unroll
    for (int wi = 0; wi < WPT; wi++) {
        const int wd = tid + wi * NT;
        const int qv = wd / S, wv = wd % S;
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
                // rotation amount of the funnel shift that swaps the lanes -- or, ARITH 11, the selector of the PRMT that does:
                // bytes (s0, 4 + s0 + 1, s2, 4 + s2 + 1) with (s0, s2) = (0, 2) or (2, 0).  With both operands the same word
                // that is the lane swap; with 0x64646464 as the second operand it also sets the fp16 exponent of 1024 + C
                if constexpr (ARITH == 11) pswp[ps][wi] = ((phi_hi ^ borrow) & 1) ? 0x5072u : 0x7250u;
                else pswp[ps][wi] = ((phi_hi ^ borrow) & 1) ? 16u : 0u;
            }
        });
    }

    uint32_t hbw[WPT];    // word index (within a column) of this warp's 32 hard bits, lane 0 / lane 1 = +S/32
#pragma unroll
    for (int wi = 0; wi < WPT; wi++) {
        const int wd = tid + wi * NT;
        hbw[wi] = (uint32_t)(((wd / S) * Q + (wd % S)) >> 5);
    }
    const uint32_t c254 = 0x00fe00feu * one, c255 = 0x00ff00ffu * one, c256 = one << 8;
    // KNOBS bit 5: in-thread row-0 exit test.  Row 0 is I(CA) + I(CP) + P(CP).  The two identity terms of a check belong
    // to the thread that owns the check; the marginal of the permuted term travels in the spare high byte of that block's
    // message (one IMAD on the variable side), so every thread tests its own row-0 checks while it processes them: no
    // ballots, no hard-bit words and no serial syndrome pass per iteration.  The marginals of all columns are kept as
    // bytes (two columns per register, one PRMT per pair) for the rare second stage and for the output.
    constexpr bool PIGGY = (KNOBS & 32) != 0;
    constexpr int NG = (NCOL + 1) / 2;     // PIGGY: registers of gathered marginal bytes per word slot
    constexpr bool CV_F = (KNOBS & 1) != 0, SUF_F = (KNOBS & 2) != 0, PRE_F = (KNOBS & 4) != 0, COMB_F = (KNOBS & 8) != 0;

    // Frames are claimed one ahead: while frame f is decoded, the LLRs of the next claimed frame are
    // already in flight (one 16-byte-aligned bulk copy of N bytes, completion on an mbarrier).
    const bool use_bulk = (reinterpret_cast<uintptr_t>(llrs_all) & 15u) == 0;   // unaligned caller buffer: plain loads
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        mbar_init_fence();
        const unsigned long long f0 = atomicAdd(counter, 1ull);
        s_frame[0] = f0;
        if (use_bulk && f0 < batch) bulk_load(stage, in_all + f0 * (unsigned long long)FB, FB, &s_bar[0]);
    }
    __syncthreads();
    unsigned cur = 0, bar_parity = 0;    // bit b of bar_parity = phase parity of s_bar[b]
    unsigned pf_var = 0, pf_bar = 0, pf_chk = 0, pf_exit = 0, pf_t = 0;
    auto pf_lap = [&](unsigned &acc) {
        if constexpr (PROF) {
            const unsigned now = (unsigned)clock();
            acc += now - pf_t;
            pf_t = now;
        }
    };

    for (;;) {
        const unsigned long long frame = s_frame[cur];
        if (frame >= batch) break;
        if (tid == 0) {                  // claim the next frame and start its copy into the other buffer
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
        uint32_t Lb[NCOL][WPT];       // channel LLR + 128 (punctured column: 128)
        uint32_t idm[NI > 0 ? NI : 1][WPT];   // identity-block messages (u after the check phase, C after the variable phase)
        uint32_t cc[NB][WPT];         // corrected C of the previous iteration (sign class of the old v)
#pragma unroll
        for (int wi = 0; wi < WPT; wi++) {
            const int wd = tid + wi * NT;
            const int e0 = (wd / S) * Q + (wd % S);
#pragma unroll
            for (int c = 0; c < NCOL; c++) {
                if (c < NCOL - 1) {
                    const int l0 = front_load<FRONT, int8_t>(llr, c * M + e0, fscale, flimit);
                    const int l1 = front_load<FRONT, int8_t>(llr, c * M + e0 + S, fscale, flimit);
                    Lb[c][wi] = (uint32_t)(l0 + 128) | ((uint32_t)(l1 + 128) << 16);
                } else {
                    Lb[c][wi] = 0x00800080u;                                      // :383
                }
            }
#pragma unroll
            for (int i = 0; i < NI; i++) idm[i][wi] = 0;
#pragma unroll
            for (int b = 0; b < NB; b++) cc[b][wi] = (ARITH >= 8 && ARITH <= 11) ? 0u : 0x007f007fu;
#pragma unroll
            for (int p = 0; p < NP; p++) msg[p * (M / 2) + wd] = 0;
        }
        for (int i = tid; i < HBW; i += NT) hb[i] = 0;
        __syncthreads();

        unsigned iters_run = max_iters;
        bool ok = false;
        uint32_t pack[WPT];           // hard bits of the columns not covered by stage 1
        uint32_t gat[PIGGY ? NG : 1][WPT];   // PIGGY: biased marginals as bytes: column 2j in bytes 0 / 2, column 2j+1 in bytes 1 / 3
        uint32_t hloc8[WPT], bad[WPT];       // PIGGY: bit 15 / 31 = XOR of the two in-thread terms; row-0 checks that failed
        bool hb_complete = true;      // hb[] holds every column's hard bits of the latest variable phase
        // ballot-packs the deferred columns' hard bits into hb[] (stage 2 / final output)
        auto flush_pack = [&]() {
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                if constexpr (PIGGY) {
                    static_for<0, NCOL>([&](auto ci) {
                        constexpr int c = decltype(ci)::value;
                        const uint32_t g = gat[c / 2][wi] >> ((c & 1) * 8);             // bit 7 / 23: marginal >= 0
                        if constexpr (S % 32 == 0) {
                            const unsigned b0 = __ballot_sync(0xFFFFFFFFu, (g & 0x00000080u) == 0);
                            const unsigned b1 = __ballot_sync(0xFFFFFFFFu, (g & 0x00800000u) == 0);
                            if (lane == 0) {
                                hb[hbw[wi] + c * M / 32] = b0;
                                hb[hbw[wi] + (c * M + S) / 32] = b1;
                            }
                        } else {        // a warp spans several half quarters: every thread ORs its two bits in (hb[] zeroed first)
                            const int wd = tid + wi * NT;
                            const int e0 = c * M + (wd / S) * Q + (wd % S);
                            if ((g & 0x00000080u) == 0) atomicOr(&hb[e0 >> 5], 1u << (e0 & 31));
                            if ((g & 0x00800000u) == 0) atomicOr(&hb[(e0 + S) >> 5], 1u << ((e0 + S) & 31));
                        }
                    });
                    continue;
                }
                int kpos = 0;
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = NCOL - 1 - decltype(ci)::value;     // last packed column sits at bit 7 / 23
                    if constexpr (c != CA && c != CP) {
                        const unsigned b0 = __ballot_sync(0xFFFFFFFFu, (pack[wi] >> (7 + kpos)) & 1u);
                        const unsigned b1 = __ballot_sync(0xFFFFFFFFu, (pack[wi] >> (23 + kpos)) & 1u);
                        if (lane == 0) {
                            hb[hbw[wi] + c * M / 32] = b0;
                            hb[hbw[wi] + (c * M + S) / 32] = b1;
                        }
                        kpos++;
                    }
                });
            }
        };
        for (unsigned iter = 0; iter < max_iters; iter++) {
            // ================= variable phase (:382-411 and :421) =================
            if constexpr (PROF) pf_t = (unsigned)clock();
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                pack[wi] = 0;
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    uint32_t va = Lb[c][wi];
                    uint32_t ub[6];
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            uint32_t u;
                            if constexpr (P::blk(b).isp) {
                                constexpr int ps = count_p<P>(b);
                                const uint32_t w = msg[paddr[ps][wi]];
                                if constexpr (ARITH == 11) u = __byte_perm(w, w, pswp[ps][wi]);
                                else u = lane_rot(w, pswp[ps][wi]);
                            } else {
                                u = idm[count_i<P>(b)][wi];
                            }
                            if constexpr (ARITH == 4 || ARITH == 7) {
                                // sign-magnitude (fp16 subnormal +-mu) -> two's complement, on the FMA pipe only:
                                // +-mu*2^-24 + 1.5*2^-14 has the bit pattern 0x0600 +- mu; then subtract 0x0600 per lane
                                u = __vadd2(h2u(__hadd2(u2h(u), u2h(0x06000600u))), 0xFA00FA00u);
                            }
                            ub[k] = u;
                            if constexpr (ARITH != 2) {
                                va = __viaddmin_s16x2_relu(va, u, 0x00ff00ffu);  // saturating_add, ascending idx
                            } else if constexpr ((k & 1) == 0) {                 // state is VA -> 255 - clamp(VA + u)
                                va = f_relu_sub(0x00ff00ffu, f_relu_add(va, u));
                            } else {                                             // state is 255 - VA -> clamp(VA + u)
                                va = f_relu_sub(0x00ff00ffu, f_relu_sub(va, u));
                            }
                        }
                    });
                    uint32_t van;
                    if constexpr (ARITH != 2) {
                        van = c255 * one - va;
                    } else if constexpr ((col_degree<P>(c) & 1) == 0) {          // chain ended in the VA form
                        van = f_sub(0x00ff00ffu, va);
                    } else {                                                     // chain ended in the 255 - VA form
                        van = va;
                        va = f_sub(0x00ff00ffu, van);
                    }
                    // hard decisions of the marginals: va < 0  <=>  VA < 128  <=>  bit 7 clear
                    if constexpr (PIGGY) {
                        if constexpr ((c & 1) == 0) gat[c / 2][wi] = va;               // bytes 1 / 3 are zero
                        else gat[c / 2][wi] = __byte_perm(gat[c / 2][wi], va, 0x6240);
                        if constexpr (c == CA) hloc8[wi] = va;
                        if constexpr (c == CP) hloc8[wi] = (hloc8[wi] ^ va) * c256;     // bit 15 / 31, no carry between the lanes
                    } else if constexpr (c == CA || c == CP) {
                        const unsigned b0 = __ballot_sync(0xFFFFFFFFu, (va & 0x00000080u) == 0);
                        const unsigned b1 = __ballot_sync(0xFFFFFFFFu, (va & 0x00800000u) == 0);
                        if (lane == 0) {
                            hb[hbw[wi] + c * M / 32] = b0;
                            hb[hbw[wi] + (c * M + S) / 32] = b1;
                        }
                    } else {
                        pack[wi] = pack[wi] * (one + one) + (~va & 0x00800080u);   // older columns move up one bit
                    }
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            // C = 127 - clamp(va - u, -127, 127)
                            uint32_t cv;
                            if constexpr (ARITH != 2) cv = __viaddmin_s16x2_relu(van, ub[k], 0x00fe00feu);
                            else cv = pmin<CV_F>(f_relu_add(van, ub[k]), 0x00fe00feu);
                            if constexpr (PIGGY && b == 2) cv = va * c256 + cv;          // high byte: the biased marginal
                            else if constexpr ((ARITH >= 8 && ARITH <= 10) || (ARITH == 11 && !P::blk(b).isp))
                                cv = cv * one + 0x64006400u;                                 // as fp16: 1024 + C
                            if constexpr (P::blk(b).isp) {
                                constexpr int ps = count_p<P>(b);
                                if constexpr (ARITH == 11 && PIGGY && b == 2) msg[paddr[ps][wi]] = __byte_perm(cv, cv, pswp[ps][wi]);
                                else if constexpr (ARITH == 11) msg[paddr[ps][wi]] = __byte_perm(cv, 0x64646464u, pswp[ps][wi]);   // swap + exponent
                                else msg[paddr[ps][wi]] = lane_rot(cv, pswp[ps][wi]);
                            } else {
                                idm[count_i<P>(b)][wi] = cv;
                            }
                        }
                    });
                });
            }
            pf_lap(pf_var);
            __syncthreads();
            pf_lap(pf_bar);

            // ================= check phase (:391-405 and :422-447) =================
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                const int wd = tid + wi * NT;
                static_for<0, NROW>([&](auto ri) {
                    constexpr int r = decltype(ri)::value;
                    constexpr int DC = row_degree<P>(r);
                    uint32_t a[kMaxDeg], ck[kMaxDeg], mu[kMaxDeg];
                    uint32_t sx = 0;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            uint32_t cv;
                            if constexpr (P::blk(b).isp) cv = msg[count_p<P>(b) * (M / 2) + wd];
                            else cv = idm[count_i<P>(b)][wi];
                            if constexpr (PIGGY && b == 2) {
                                // three marginals of this check: (biased bit 7 of each) XORed = NOT the parity of the hard bits
                                bad[wi] = ~(hloc8[wi] ^ cv) & 0x80008000u;
                                if constexpr (ARITH >= 8 && ARITH <= 11) cv = (cv & 0x00ff00ffu) | 0x64006400u;
                                else cv &= 0x00ff00ffu;
                            }
                            if constexpr (ARITH >= 8 && ARITH <= 11) {
                                const __half2 d = __hsub2(u2h(0x647f647fu), u2h(cv));      // 1151 - (1024 + C) = v, an integer-valued fp16 (0 -> +0)
                                const __half2 keep = __hfma2_sat(d, u2h(cc[b][wi]), u2h(0x3c003c00u));   // sat(v v_old + 1): 0 where the sign flipped and v_old != 0
                                const __half2 dc = __hfma2(d, keep, u2h(0u));              // killed -> +0
                                cc[b][wi] = h2u(dc);
                                ck[k] = h2u(dc);
                                a[k] = h2u(dc);
                                sx ^= h2u(dc);                                             // bit 15: product of signs
                            } else {
                            const uint32_t old = cc[b][wi];
                            const uint32_t x = (cv ^ old) & (cv ^ (old + 0x00010001u));   // bit 7: sign flipped and old != 0
                            const uint32_t km = prmt_sign7(x);
                            const uint32_t cor = (cv & ~km) | (0x007f007fu & km);         // killed -> v = 0
                            cc[b][wi] = cor;
                            ck[k] = cor;
                            if constexpr (ARITH == 1) a[k] = __vmaxu2(cor, c254 * one - cor);    // |v| + 127
                            else if constexpr (ARITH == 6 || ARITH == 7)
                                a[k] = h2u(__hsub2(u2h(cor), u2h(0x007f007fu)));                 // -v as a signed fp16 lane
                            else a[k] = __vabsdiffu4(cor, 0x007f007fu);                          // |v|
                            sx ^= cor;                                                     // bit 7: product of signs
                            }
                        }
                    });
                    if constexpr (ARITH == 9) sx &= 0x80008000u;                          // +-0 in both lanes
                    if constexpr (ARITH >= 10) sx = (sx & 0x80008000u) ^ 0x3c003c00u;     // +-1.0 in both lanes
                    if constexpr (ARITH == 5) min_excluding_self3<DC>(a, mu);
                    else if constexpr (ARITH >= 6 && ARITH <= 11) min_excluding_self_h<DC, (KNOBS & 16) != 0>(a, mu);
                    else if constexpr (ARITH == 1 || ARITH == 3) min_excluding_self<DC, false, false, false>(a, mu);
                    else min_excluding_self<DC, SUF_F, PRE_F, COMB_F>(a, mu);
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            uint32_t u;
                            if constexpr (ARITH == 1) {
                                const uint32_t nm = prmt_sign7(sx ^ ck[k]);                // lanes whose u is negative
                                const uint32_t kk = __vadd2(nm, 0xff81ff81u);              // -127 or -128
                                u = __vadd2(mu[k], kk) ^ nm;                               // +-(mu - 127), two's complement
                            } else if constexpr (ARITH == 3 || ARITH == 5 || ARITH == 6) {
                                const uint32_t nm = prmt_sign7(sx ^ ck[k]);                // lanes whose u is negative
                                u = __vadd2(mu[k], nm) ^ nm;                               // +-mu, two's complement
                            } else if constexpr (ARITH == 8) {
                                const uint32_t nm = prmt_sign15(sx ^ ck[k]);
                                const uint32_t mi = h2u(__hmul2(u2h(mu[k]), u2h(0x00010001u)));   // mu * 2^-24: the integer in the low bits
                                u = __vadd2(mi, nm) ^ nm;
                            } else if constexpr (ARITH >= 10) {
                                // +-1.0 with the sign of u; 1536 +- mu has the bits 0x6600 +- mu (ulp 1 in [1024, 2048))
                                const uint32_t pm = sx ^ (ck[k] & 0x80008000u);
                                u = __vadd2(h2u(__hfma2(u2h(mu[k]), u2h(pm), u2h(0x66006600u))), 0x9a009a00u);
                            } else if constexpr (ARITH == 9) {
                                const uint32_t nm = prmt_sign15(h2u(__hmul2(u2h(sx), u2h(ck[k]))));   // the sign of a product survives a zero factor
                                const uint32_t mi = h2u(__hmul2(u2h(mu[k]), u2h(0x00010001u)));
                                u = __vadd2(mi, nm) ^ nm;
                            } else {
                                const uint32_t zs = (sx ^ ck[k]) & 0x00800080u;            // sign of u at bit 7
                                u = zs * c256 + mu[k];                                     // sign-magnitude: bit 15 | mu
                            }
                            if constexpr (P::blk(b).isp) msg[count_p<P>(b) * (M / 2) + wd] = u;
                            else idm[count_i<P>(b)][wi] = u;
                        }
                    });
                });
            }
            pf_lap(pf_chk);
            // ---- parity of the marginals' hard bits (:445-453), one thread per 32 checks ----
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
            // stage 1: row 0 only (its hard bits were packed in this iteration's variable phase)
            uint32_t synd = 0;
            if constexpr (PIGGY) {
#pragma unroll
                for (int wi = 0; wi < WPT; wi++) synd |= bad[wi];
            } else {
                // (done by the HIGHEST-numbered warps: the SM's arbiter favours high warp ids, so the warps with the
                //  extra work are the ones that reach the barrier early anyway)
                if (tid >= NT - M / 32) synd = syndrome_word(tid - (NT - M / 32));
            }
            hb_complete = false;
            if (__syncthreads_or(synd != 0) == 0) {
                // stage 2: row 0 is clean -- pack the other columns and test rows 1..NROW-1
                if constexpr (PIGGY && S % 32 != 0) {
                    for (int i = tid; i < HBW; i += NT) hb[i] = 0;
                    __syncthreads();
                }
                flush_pack();
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
            pf_lap(pf_exit);
        }
        if (!hb_complete) {       // decoding failed: the output is the hard decision of the last marginals (:466-473)
            if constexpr (PIGGY && S % 32 != 0) {
                for (int i = tid; i < HBW; i += NT) hb[i] = 0;
                __syncthreads();
            }
            flush_pack();
            __syncthreads();
        }

        // ---- output: hard decisions of all n+p marginals, MSB first (:455-461, :466-473) ----
        uint8_t *out = out_all + frame * (unsigned long long)(NV / 8);
        const bool aligned = (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
        for (int i = tid; i < HBW; i += NT) {
            const uint32_t rev = __brev(hb[i]);                    // variable 32i at bit 31
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
    if constexpr (PROF) {
        if (lane == 0) {
            unsigned long long *pw = prof + (tid >> 5) * 4;
            atomicAdd(pw + 0, (unsigned long long)pf_var);
            atomicAdd(pw + 1, (unsigned long long)pf_bar);
            atomicAdd(pw + 2, (unsigned long long)pf_chk);
            atomicAdd(pw + 3, (unsigned long long)pf_exit);
        }
    }
}

template <int RATE, int M, int WPT, int ARITH = 2, int KNOBS = 2 + 1, int MINB = 1, int FRONT = kFrontNone, bool PROF = false>
cudaError_t launch_tm(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                      size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                      const Front &front = Front()) {
    typedef Proto<RATE> P;
    constexpr int NP = count_p<P>(P::NB);
    constexpr int NT = M / 2 / WPT;
    const TmParams prm = make_params<RATE>(c);
    const size_t smem = ((size_t)NP * (M / 2) + (((size_t)P::NCOL * M / 32 + 3) & ~(size_t)3)) * sizeof(uint32_t) +
                        2 * front_frame_bytes(front, (P::NCOL - 1) * M, kI8);   // messages + hard bits + two input staging buffers
    auto kern = decode_ms_tm_i8_kernel<RATE, M, WPT, ARITH, KNOBS, MINB, FRONT, PROF>;
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
    static const int cap = [] { const char *e = getenv("LABRADOR_LDPC_TM_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
    if (cap > 0 && per_sm > cap) per_sm = cap;     // A/B runs: fewer resident CTAs than fit
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm;
    if (grid > batch) grid = batch;
    WorkCounter wc(ctx, stream);
    if (wc.error() != cudaSuccess) return wc.error();
    unsigned long long *counter = wc.ptr();
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    unsigned long long *prof = nullptr;
    if constexpr (PROF) {
        e = cudaMalloc(&prof, 32 * 4 * sizeof(unsigned long long));
        if (e != cudaSuccess) return e;
        cudaMemsetAsync(prof, 0, 32 * 4 * sizeof(unsigned long long), stream);
    }
    kern<<<(unsigned)grid, NT, smem, stream>>>(prm, static_cast<const typename FrontSrc<FRONT, int8_t>::type *>(llrs), output,
                                               (unsigned long long)batch, mi, success, iters, counter, 1u, front.scale,
                                               front.limit, prof);
    count_launch();
    if constexpr (PROF) {     // development aid: per-warp cycle totals of the four parts of an iteration, to stderr
        unsigned long long h[32 * 4];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(prof);
        fprintf(stderr, "tm_prof arith=%d knobs=%d grid=%llu threads=%d batch=%zu\n", ARITH, KNOBS, grid, NT, batch);
        for (int w = 0; w < NT / 32; w++)
            fprintf(stderr, "tm_prof warp %2d var %llu bar %llu chk %llu exit %llu\n", w, h[w * 4], h[w * 4 + 1], h[w * 4 + 2],
                    h[w * 4 + 3]);
    }
    return cudaGetLastError();
}

}  // namespace

// Variant selection.  Several arithmetic variants are compiled per code; the default (1032 = ARITH 10 with the in-thread
// exit test, KNOBS 32) is the one that measured fastest on B200 for every code (profiles/r02_tm_variants.md).
// LABRADOR_LDPC_TM_ARITH=1|2|3|4|5|6|7|532|616|632|716|832|932|1032 overrides (A/B runs); 52 / 6322 / 9322 / 10322
// (TM5120 only) = ARITH 5 / 632 / 932 / 1032 compiled for 2 resident CTAs per SM (128 registers).
template <int RATE, int M>
cudaError_t launch_tm_variant(int default_arith, DeviceCtx &ctx, const CodeInfo &c, const void *l, uint8_t *output,
                              size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    static const int forced = [] { const char *e = getenv("LABRADOR_LDPC_TM_ARITH"); return e ? atoi(e) : 0; }();
    const int arith = forced ? forced : default_arith;
    if constexpr (M == 2048) {   // TM8192: one word slot per thread (1024 threads x 64 registers, default: +1.4 % with the fp16
                                 // check side, it was -3.2 % with the integer one) or two (512 threads x <= 128 registers:
                                 // LABRADOR_LDPC_TM_WPT=2, the fused front ends and the cycle profile of LABRADOR_LDPC_TM_PROF)
        static const bool prof = [] { const char *e = getenv("LABRADOR_LDPC_TM_PROF"); return e && atoi(e) != 0; }();
        static const int wpt = [] { const char *e = getenv("LABRADOR_LDPC_TM_WPT"); return e ? atoi(e) : (prof ? 2 : 1); }();
        if (wpt == 2) {
            if (prof) {
                if (arith == 2) return launch_tm<RATE, M, 2, 2, 6, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                if (arith == 4) return launch_tm<RATE, M, 2, 4, 0, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                if (arith == 6) return launch_tm<RATE, M, 2, 6, 0, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                if (arith == 7) return launch_tm<RATE, M, 2, 7, 0, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                if (arith == 632) return launch_tm<RATE, M, 2, 6, 32, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                if (arith == 932) return launch_tm<RATE, M, 2, 9, 32, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                if (arith == 1032) return launch_tm<RATE, M, 2, 10, 32, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                return launch_tm<RATE, M, 2, 5, 0, 1, kFrontNone, true>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            }
            if (arith == 832) return launch_tm<RATE, M, 2, 8, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 932) return launch_tm<RATE, M, 2, 9, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 1032) return launch_tm<RATE, M, 2, 10, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 1132) return launch_tm<RATE, M, 2, 11, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 6) return launch_tm<RATE, M, 2, 6, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 632) return launch_tm<RATE, M, 2, 6, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 532) return launch_tm<RATE, M, 2, 5, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 616) return launch_tm<RATE, M, 2, 6, 16>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 7) return launch_tm<RATE, M, 2, 7, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 716) return launch_tm<RATE, M, 2, 7, 16>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 3) return launch_tm<RATE, M, 2, 3, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 5) return launch_tm<RATE, M, 2, 5, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            if (arith == 4) return launch_tm<RATE, M, 2, 4, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            return launch_tm<RATE, M, 2, 2, 6>(ctx, c, l, output, batch, max_iters, success, iters, stream);
        }
    }
    if (arith == 1) return launch_tm<RATE, M, 1, 1, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 3) return launch_tm<RATE, M, 1, 3, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 5) return launch_tm<RATE, M, 1, 5, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 6) return launch_tm<RATE, M, 1, 6, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 632) return launch_tm<RATE, M, 1, 6, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 832) return launch_tm<RATE, M, 1, 8, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 932) return launch_tm<RATE, M, 1, 9, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 1032) return launch_tm<RATE, M, 1, 10, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 1132) return launch_tm<RATE, M, 1, 11, 32>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if (arith == 7) return launch_tm<RATE, M, 1, 7, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    if constexpr (RATE == 2 && M == 512) {
        if (arith == 11322) return launch_tm<RATE, M, 1, 11, 32, 2>(ctx, c, l, output, batch, max_iters, success, iters, stream);
        if (arith == 10322) return launch_tm<RATE, M, 1, 10, 32, 2>(ctx, c, l, output, batch, max_iters, success, iters, stream);
        if (arith == 9322) return launch_tm<RATE, M, 1, 9, 32, 2>(ctx, c, l, output, batch, max_iters, success, iters, stream);
        if (arith == 6322) return launch_tm<RATE, M, 1, 6, 32, 2>(ctx, c, l, output, batch, max_iters, success, iters, stream);
        if (arith == 52) return launch_tm<RATE, M, 1, 5, 0, 2>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    }
    if (arith == 4) return launch_tm<RATE, M, 1, 4, 0>(ctx, c, l, output, batch, max_iters, success, iters, stream);
    return launch_tm<RATE, M, 1, 2, 6>(ctx, c, l, output, batch, max_iters, success, iters, stream);
}

// The fused front ends (front.cuh) are compiled for each code's default variant only.
template <int RATE, int M, int WPT, int ARITH, int KNOBS, int MINB>
cudaError_t launch_tm_front(const Front &front, DeviceCtx &ctx, const CodeInfo &c, const void *l, uint8_t *output,
                            size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    if (front.kind == kFrontSoftF32)
        return launch_tm<RATE, M, WPT, ARITH, KNOBS, MINB, kFrontSoftF32>(ctx, c, l, output, batch, max_iters, success,
                                                                           iters, stream, front);
    if (front.kind == kFrontHard)
        return launch_tm<RATE, M, WPT, ARITH, KNOBS, MINB, kFrontHard>(ctx, c, l, output, batch, max_iters, success, iters,
                                                                        stream, front);
    return cudaErrorInvalidValue;
}

bool has_decode_ms_tm_i8(int code);

// Returns true (and launches) if a specialised kernel exists for (code, i8).
bool launch_decode_ms_tm_i8(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch,
                            size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                            cudaError_t *err, const Front &front) {
    if (!has_decode_ms_tm_i8(code)) return false;
    const CodeInfo &c = *code_info(code);
    const void *l = llrs;
    const bool ff = front.kind != kFrontNone;
    switch (code) {
        case 3:      // TM1280: M = 128, 64 threads per codeword; needs the in-thread exit test (S = 16 < one warp)
            if (!structure_matches<2>(c) || c.m != 128) return false;
            if (ff) *err = launch_tm_front<2, 128, 1, 10, 32, 1>(front, ctx, c, l, output, batch, max_iters, success, iters, stream);
            else {
                // resident CTAs per SM the kernel is compiled for (register budget 65536 / 64 / MINB): LABRADOR_LDPC_TM1280_MINB
                static const int minb = [] { const char *e = getenv("LABRADOR_LDPC_TM1280_MINB"); return e ? atoi(e) : 6; }();
                static const int arith = [] { const char *e = getenv("LABRADOR_LDPC_TM_ARITH"); return e ? atoi(e) : 1032; }();
                if (arith == 632) *err = launch_tm<2, 128, 1, 6, 32, 6>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                else if (minb <= 1) *err = launch_tm<2, 128, 1, 10, 32, 1>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                else if (minb <= 6) *err = launch_tm<2, 128, 1, 10, 32, 6>(ctx, c, l, output, batch, max_iters, success, iters, stream);
                else *err = launch_tm<2, 128, 1, 10, 32, 8>(ctx, c, l, output, batch, max_iters, success, iters, stream);
            }
            return true;
        case 4:
            if (!structure_matches<1>(c) || c.m != 256) return false;
            if (ff) *err = launch_tm_front<1, 256, 1, 10, 32, 1>(front, ctx, c, l, output, batch, max_iters, success, iters, stream);
            else *err = launch_tm_variant<1, 256>(1032, ctx, c, l, output, batch, max_iters, success, iters, stream);
            return true;
        case 5:
            if (!structure_matches<0>(c) || c.m != 512) return false;
            if (ff) *err = launch_tm_front<0, 512, 1, 10, 32, 1>(front, ctx, c, l, output, batch, max_iters, success, iters, stream);
            else *err = launch_tm_variant<0, 512>(1032, ctx, c, l, output, batch, max_iters, success, iters, stream);
            return true;
        case 6:
            if (!structure_matches<2>(c) || c.m != 512) return false;
            if (ff) *err = launch_tm_front<2, 512, 1, 10, 32, 2>(front, ctx, c, l, output, batch, max_iters, success, iters, stream);
            else *err = launch_tm_variant<2, 512>(10322, ctx, c, l, output, batch, max_iters, success, iters, stream);
            return true;
        case 7:
            if (!structure_matches<1>(c) || c.m != 1024) return false;
            if (ff) *err = launch_tm_front<1, 1024, 1, 10, 32, 1>(front, ctx, c, l, output, batch, max_iters, success, iters, stream);
            else *err = launch_tm_variant<1, 1024>(1032, ctx, c, l, output, batch, max_iters, success, iters, stream);
            return true;
        case 8:
            if (!structure_matches<0>(c) || c.m != 2048) return false;
            if (ff) *err = launch_tm_front<0, 2048, 2, 10, 32, 1>(front, ctx, c, l, output, batch, max_iters, success, iters, stream);
            else *err = launch_tm_variant<0, 2048>(1032, ctx, c, l, output, batch, max_iters, success, iters, stream);
            return true;
        default:
            return false;
    }
}

bool has_decode_ms_tm_i8(int code) {
    static const bool tm1280_wide = [] { const char *e = getenv("LABRADOR_LDPC_TM1280_WIDE"); return e && atoi(e) != 0; }();
    return code >= (tm1280_wide ? 4 : 3) && code <= 8;      // LABRADOR_LDPC_TM1280_WIDE=1: TM1280 i8 stays on the scalar-lane kernel (A/B)
}

}  // namespace ldpc
