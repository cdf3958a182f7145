// Min-sum decoder for the k = 16384 TM codes with i8 LLRs: one codeword per CLUSTER of four CTAs.
//
// Replaces LDPCCode::decode_ms::<i8> (reference src/decoder.rs:347-475) for TM20480 (r = 4/5, M = 2048), TM24576
// (r = 2/3, M = 4096) and TM32768 (r = 1/2, M = 8192) -- codes the reference carries constants for but does not
// support (src/lib.rs:81-83); results are bit-identical to the oracle (the reference's algorithm over these codes).
//
// The packed i8 kernel (decode_ms_tm.cu) keeps a codeword's state in the registers of ONE CTA: 43 registers per word
// slot.  M = 8192 has 4096 word slots -- four register files.  The pi_k permutation of every block maps quarter q of
// the checks to quarter (theta + q) mod 4 of the variables (reference src/codes/mod.rs:312-322), so the codeword splits
// naturally four ways: CTA r of the cluster owns quarter r of every prototype column AND of every prototype row
// (M/8 word slots, each two 16-bit lanes: elements x and x + M/8 of the quarter).  Then
//   * identity blocks connect a check and a variable of the same thread (registers), as before;
//   * a permutation block connects variable quarter r with check quarter (r - theta) mod 4.  Its messages cross the
//     cluster through DISTRIBUTED SHARED MEMORY, and always as a PUSH: the variable side stores v into a buffer of the CTA
//     that owns the checks (check order), the check side stores u into a buffer of the CTA that owns the variables
//     (variable order), both with st.shared::cluster to a per-thread constant address (mapa); every load is local.
//     Distributed shared memory moves ~20 bytes per clock per SM at ~215 cycles latency (B300_MICROARCH.md): pulled at
//     the start of a phase that is 2500 exposed cycles per iteration (the first version of this kernel, 0.37 M cw/s on
//     TM32768); pushed as soon as produced, the transfer drains behind the rest of the phase's arithmetic;
//   * the two barriers of an iteration become cluster barriers (barrier.cluster arrive.release / wait.acquire, which
//     also order the remote stores), and the exit decision is an OR over the four CTAs (one flag word per CTA, written
//     into every CTA's shared memory before the barrier that the next phase needs anyway).
// Two kernels follow.  decode_ms_tm_cluster_kernel is the form described so far (two messages per 32-bit word, the
// integer self-correction rule: ARITH 6 + KNOBS 32 of decode_ms_tm.cu), kept for A/B runs and tests
// (LABRADOR_LDPC_CLUSTER_QUAD=0).  decode_ms_tm_cluster4_kernel is the one that ships: FOUR messages per word (half the
// stores and bytes across the cluster), completion of a phase by BYTE COUNT -- st.async stores that complete on an
// mbarrier of the receiving CTA, so an iteration contains no cluster barrier -- and the check side as exact fp16
// arithmetic (ARITH 10 of decode_ms_tm.cu).  Measurements of every step: profiles/r02_cluster.md.
// In both: biased s16x2 lanes, VIADDMNMX saturating adds, fp16 |v| and minima, row-0 exit test in the threads that own
// the checks with the permuted marginal riding in the message's high byte; rows 1-2 only when row 0 is clean, from
// ballot-packed hard bits (the other quarters' words are read through distributed shared memory).  Each CTA stages its
// quarter of the next frame (one bulk asynchronous copy per data column) while the current frame is decoded.
// Frames are assigned statically (cluster c takes frames c, c + #clusters, ...): a claim would have to be broadcast.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "bulk_copy.cuh"
#include "runtime.h"
#include "tm_common.cuh"

namespace ldpc {
using namespace tm;

namespace {

constexpr int kCL = 4;           // CTAs per cluster = quarters of a block
constexpr int kMaxDegC = 18;

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_count() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA's window) in CTA `rank`
__device__ __forceinline__ uint32_t cl_map(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cl_ld(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cl_st(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Remote store that completes on an mbarrier of the DESTINATION CTA (st.async, sm_90+): the receiver waits for a byte
// count instead of a cluster barrier, the sender needs no release fence.
__device__ __forceinline__ void cl_st_async(uint32_t addr, uint32_t v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t bar_sa, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_sa), "r"(bytes) : "memory");
}
// wait for a phase completed by other CTAs' st.async (acquire at cluster scope).  A protocol error would spin for
// ever: after 2^26 polls (seconds; a phase takes microseconds) the kernel traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar_sa, uint32_t parity) {
    uint32_t done = 0, polls = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar_sa), "r"(parity) : "memory");
        if (done) break;
        if (++polls == (1u << 26)) __trap();
    }
}

__device__ __forceinline__ uint32_t sign7_mask(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xaa88;" : "=r"(r) : "r"(x));
    return r;
}
// bytes 0, 2 (1, 3) of x sign-extended into the two 16-bit lanes
__device__ __forceinline__ uint32_t sext_bytes02(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xa280;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t sext_bytes13(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xb391;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t lrot(uint32_t x, uint32_t sh) { return __funnelshift_l(x, x, sh); }
__device__ __forceinline__ __half2 cu2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t ch2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint32_t chmin2a(uint32_t x, uint32_t y) { return ch2u(__hmin2(__habs2(cu2h(x)), __habs2(cu2h(y)))); }
__device__ __forceinline__ uint32_t chmin3a(uint32_t x, uint32_t y, uint32_t z) {
    return ch2u(__hmin2(__hmin2(__habs2(cu2h(x)), __habs2(cu2h(y))), __habs2(cu2h(z))));
}
// minimum of |.| over the other edges of one check word (fp16 lanes, see decode_ms_tm.cu: min_excluding_self_h)
template <int DC>
__device__ __forceinline__ void min_excl_h(const uint32_t (&a)[kMaxDegC], uint32_t (&mu)[kMaxDegC]) {
    static_assert(DC >= 3, "every output must come out of a minimum");
    constexpr int NPAIR = DC / 2;
    constexpr bool ODD = (DC & 1) != 0;
    uint32_t suf[kMaxDegC / 2 + 2];
    if constexpr (ODD) suf[NPAIR] = a[DC - 1];
#pragma unroll
    for (int j = NPAIR - 1; j >= 1; j--) {
        if (j == NPAIR - 1 && !ODD) suf[j] = chmin2a(a[2 * j], a[2 * j + 1]);
        else suf[j] = chmin3a(a[2 * j], a[2 * j + 1], suf[j + 1]);
    }
    uint32_t pre = 0;
#pragma unroll
    for (int j = 0; j < NPAIR; j++) {
        const bool has_pre = j > 0, has_suf = (j + 1 < NPAIR) || ODD;
        if (has_pre && has_suf) {
            mu[2 * j] = chmin3a(pre, a[2 * j + 1], suf[j + 1]);
            mu[2 * j + 1] = chmin3a(pre, a[2 * j], suf[j + 1]);
        } else if (has_suf) {
            mu[2 * j] = chmin2a(a[2 * j + 1], suf[j + 1]);
            mu[2 * j + 1] = chmin2a(a[2 * j], suf[j + 1]);
        } else {
            mu[2 * j] = chmin2a(pre, a[2 * j + 1]);
            mu[2 * j + 1] = chmin2a(pre, a[2 * j]);
        }
        if (j + 1 < NPAIR || ODD) pre = has_pre ? chmin3a(pre, a[2 * j], a[2 * j + 1]) : chmin2a(a[2 * j], a[2 * j + 1]);
    }
    if constexpr (ODD) mu[DC - 1] = pre;
}

template <int RATE, int M, int WPT, int MINB, bool ASYNC>
__global__ void __cluster_dims__(kCL, 1, 1) __launch_bounds__(M / 8 / WPT, MINB)
decode_ms_tm_cluster_kernel(const TmParams prm, const int8_t *__restrict__ llrs_all, uint8_t *__restrict__ out_all,
                            unsigned long long batch, unsigned max_iters, uint8_t *__restrict__ success,
                            uint32_t *__restrict__ iters_out,
                            const uint32_t one /* == 1: keeps carry-free packing and subtractions on the FMA pipe (IMAD) */) {
    typedef Proto<RATE> P;
    constexpr int NB = P::NB, NCOL = P::NCOL, NROW = P::NROW;
    constexpr int NP = count_p<P>(NB), NI = NB - NP;
    constexpr int Q = M / 4, S = Q / 2, NT = S / WPT;          // S word slots per CTA (= per quarter)
    constexpr int NV = NCOL * M, N = (NCOL - 1) * M;
    constexpr int QW = Q / 32;                                  // hard-decision words per column per quarter
    constexpr int HBL = NCOL * QW;                              // ... per CTA
    constexpr int CA = P::blk(0).col, CP = NCOL - 1;
    constexpr int NG = (NCOL + 1) / 2;
    static_assert(S % 32 == 0 && NT % 32 == 0 && S % NT == 0, "whole warps per half quarter");
    static_assert(P::blk(0).row == 0 && !P::blk(0).isp && P::blk(1).row == 0 && P::blk(1).col == CP && !P::blk(1).isp &&
                  P::blk(2).row == 0 && P::blk(2).col == CP && P::blk(2).isp && P::blk(3).row == 1,
                  "row 0 must be I(CA) + I(CP) + P(CP)");

    extern __shared__ __align__(16) uint32_t smem_cl[];
    uint32_t *msg = smem_cl;                         // [NP][S] v messages of this CTA's CHECK quarter, check order (pushed by the variable side)
    uint32_t *ubuf = msg + NP * S;                   // [NP][S] u messages of this CTA's VARIABLE quarter, variable order (pushed by the check side)
    uint2 *tab = reinterpret_cast<uint2 *>(ubuf + NP * S);   // [NP][WPT][NT] check side: {shared::cluster address in ubuf, lane swap}
    uint32_t *hb = reinterpret_cast<uint32_t *>(tab + NP * WPT * NT);   // [NCOL][QW] packed hard decisions of this CTA's variable quarter
    constexpr unsigned FBL = (NCOL - 1) * Q;         // staged bytes per frame: this quarter of every data column
    unsigned char *stage = reinterpret_cast<unsigned char *>(hb + ((HBL + 3) & ~3));   // [2][FBL]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_flag[kCL];                 // exit-test flags of the four CTAs (written by their thread 0)
    // ASYNC: [0] counts the bytes of v pushed into msg, [1] the bytes of u pushed into ubuf plus the four row-0 flags
    __shared__ __align__(8) uint64_t s_abar[2];
    __shared__ uint32_t s_vbar[NP > 0 ? NP : 1], s_ubar[NP > 0 ? NP : 1];   // per block: the destination CTA's s_abar[0] / [1]
    __shared__ uint32_t s_aflag[2][kCL];             // row-0 flags by iteration parity
    constexpr uint32_t kPhaseBytes = (uint32_t)NP * S * 4u;

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t rank = cl_rank();                 // = the quarter this CTA owns
    const uint32_t msg_sa = smem_addr(msg), ubuf_sa = smem_addr(ubuf), hb_sa = smem_addr(hb), flag_sa = smem_addr(s_flag);
    const uint32_t vbar_sa = smem_addr(&s_abar[0]), ubar_sa = smem_addr(&s_abar[1]), aflag_sa = smem_addr(&s_aflag[0][0]);

    // per-thread constants: shared::cluster address + lane swap of every permutation block (variable side)
    uint32_t paddr[NP > 0 ? NP : 1][WPT], pswp[NP > 0 ? NP : 1][WPT];
#pragma unroll
    for (int wi = 0; wi < WPT; wi++) {
        const int wv = tid + wi * NT;
        static_for<0, NB>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            if constexpr (P::blk(b).isp) {
                constexpr int ps = count_p<P>(b);
                const int q = ((int)rank - (int)prm.theta[b]) & 3;           // the check quarter = the CTA that holds the message
                const int phi = prm.phi[b][q];
                const int phi_lo = phi % S, phi_hi = phi / S;
                const int borrow = wv < phi_lo ? 1 : 0;
                const int w = (wv - phi_lo) & (S - 1);
                paddr[ps][wi] = cl_map(msg_sa + (uint32_t)(ps * S + w) * 4u, (uint32_t)q);
                pswp[ps][wi] = ((phi_hi ^ borrow) & 1) ? 16u : 0u;
                // the same block seen from the check this thread owns (quarter `rank`, slot wv): the variable pair it talks to
                const int qv2 = ((int)prm.theta[b] + (int)rank) & 3;
                const int phi2 = prm.phi[b][rank];
                const int t2 = wv + phi2 % S;                                 // carry out of the half quarter <=> the var side's borrow
                const int wv2 = t2 & (S - 1);
                tab[(ps * WPT + wi) * NT + tid] = make_uint2(cl_map(ubuf_sa + (uint32_t)(ps * S + wv2) * 4u, (uint32_t)qv2),
                                                             (((phi2 / S) ^ (t2 >= S ? 1 : 0)) & 1) ? 16u : 0u);
                if (ASYNC && tid == 0 && wi == 0) {
                    s_vbar[ps] = cl_map(vbar_sa, (uint32_t)q);
                    s_ubar[ps] = cl_map(ubar_sa, (uint32_t)qv2);
                }
            }
        });
    }
    const uint32_t c255 = 0x00ff00ffu * one, c256 = one << 8;

    const unsigned long long n_clusters = cl_count();
    const bool use_bulk = (reinterpret_cast<uintptr_t>(llrs_all) & 15u) == 0;
    // one thread: this quarter of every data column of `frame` into staging buffer `buf` (NCOL-1 bulk copies, one barrier)
    auto stage_frame = [&](unsigned long long frame, unsigned buf) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_bar[buf])), "r"(FBL) : "memory");
        const int8_t *src = llrs_all + frame * (unsigned long long)N + rank * Q;
#pragma unroll 1
        for (int c = 0; c < NCOL - 1; c++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(stage + buf * FBL + c * Q)), "l"(src + (size_t)c * M), "r"((unsigned)Q),
                           "r"(smem_addr(&s_bar[buf])) : "memory");
    };
    unsigned long long frame = cl_id();
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        if (ASYNC) {
            mbar_init(&s_abar[0], 1);
            mbar_init(&s_abar[1], 1);
        }
        mbar_init_fence();
        if (use_bulk && frame < batch) stage_frame(frame, 0);
    }
    __syncthreads();
    if (ASYNC) cl_sync();                // every CTA's barriers exist before anyone's stores complete on them
    unsigned cur = 0, bar_parity = 0;
    uint32_t apar = 0;                   // ASYNC: parity of the next phase of s_abar[0] (bit 0) / s_abar[1] (bit 1); both flip once per iteration

    for (; frame < batch; frame += n_clusters) {
        if (tid == 0 && use_bulk && frame + n_clusters < batch) stage_frame(frame + n_clusters, cur ^ 1);
        if (use_bulk) {
            mbar_wait(&s_bar[cur], (bar_parity >> cur) & 1u);
            bar_parity ^= 1u << cur;
        }
        const int8_t *llr_s = reinterpret_cast<const int8_t *>(stage + cur * FBL);                    // [NCOL-1][Q]
        const int8_t *llr_g = llrs_all + frame * (unsigned long long)N + rank * Q;                     // column stride M

        // ---- per-frame state: everything zero, every call (:368, :374) ----
        uint32_t Lb[NCOL][WPT], idm[NI > 0 ? NI : 1][WPT], cc[NB][WPT];
#pragma unroll
        for (int wi = 0; wi < WPT; wi++) {
            const int wv = tid + wi * NT;
#pragma unroll
            for (int c = 0; c < NCOL; c++) {
                if (c < NCOL - 1) {
                    const int l0 = use_bulk ? llr_s[c * Q + wv] : llr_g[(size_t)c * M + wv];
                    const int l1 = use_bulk ? llr_s[c * Q + wv + S] : llr_g[(size_t)c * M + wv + S];
                    Lb[c][wi] = (uint32_t)(l0 + 128) | ((uint32_t)(l1 + 128) << 16);
                } else {
                    Lb[c][wi] = 0x00800080u;                                      // :383
                }
            }
#pragma unroll
            for (int i = 0; i < NI; i++) idm[i][wi] = 0;
#pragma unroll
            for (int b = 0; b < NB; b++) cc[b][wi] = 0x007f007fu;
#pragma unroll
            for (int p = 0; p < NP; p++) ubuf[p * S + wv] = 0;          // u = 0 before the first iteration; v is written before it is read
        }
        for (int i = tid; i < HBL; i += NT) hb[i] = 0;
        if (ASYNC) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cl_sync();                       // every CTA has re-initialised its buffers before anyone pushes into them

        unsigned iters_run = max_iters;
        bool ok = false, hb_complete = true;
        uint32_t gat[NG][WPT], hloc8[WPT], bad[WPT];
        // OR of `pred` over all threads of the four CTAs.  The flags are read right after the cluster barrier and the
        // next call writes them only after another cluster barrier, so one set of flag words suffices.
        auto cluster_or = [&](bool pred) {
            const int local = __syncthreads_or(pred);
            if (tid < kCL) cl_st(cl_map(flag_sa + rank * 4u, (uint32_t)tid), (uint32_t)local);
            cl_sync();
            return (s_flag[0] | s_flag[1] | s_flag[2] | s_flag[3]) != 0;
        };
        auto flush_pack = [&]() {        // hard bits of every column of this quarter into hb[] (second stage / output)
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                const int w32 = (tid + wi * NT) >> 5;
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    const uint32_t g = gat[c / 2][wi] >> ((c & 1) * 8);             // bit 7 / 23: marginal >= 0
                    const unsigned b0 = __ballot_sync(0xFFFFFFFFu, (g & 0x00000080u) == 0);
                    const unsigned b1 = __ballot_sync(0xFFFFFFFFu, (g & 0x00800000u) == 0);
                    if (lane == 0) {
                        hb[c * QW + w32] = b0;
                        hb[c * QW + w32 + S / 32] = b1;
                    }
                });
            }
        };

        for (unsigned iter = 0; iter < max_iters; iter++) {
            if (ASYNC && tid == 0) {     // what this CTA will receive in this iteration (stores that arrive earlier count too)
                mbar_arm(vbar_sa, kPhaseBytes);
                mbar_arm(ubar_sa, kPhaseBytes + 4u * kCL);
            }
            // ================= variable phase (:382-411 and :421) =================
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
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
                                u = ubuf[ps * S + tid + wi * NT];            // pushed here, in this thread's lane order, by the check side
                            } else {
                                u = idm[count_i<P>(b)][wi];
                            }
                            ub[k] = u;
                            va = __viaddmin_s16x2_relu(va, u, 0x00ff00ffu);      // saturating_add, ascending idx (:408)
                        }
                    });
                    const uint32_t van = c255 * one - va;
                    if constexpr ((c & 1) == 0) gat[c / 2][wi] = va;
                    else gat[c / 2][wi] = __byte_perm(gat[c / 2][wi], va, 0x6240);
                    if constexpr (c == CA) hloc8[wi] = va;
                    if constexpr (c == CP) hloc8[wi] = (hloc8[wi] ^ va) * c256;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            uint32_t cv = __viaddmin_s16x2_relu(van, ub[k], 0x00fe00feu);   // C = 127 - clamp(va - u, +-127)
                            if constexpr (b == 2) cv = va * c256 + cv;                      // high byte: the biased marginal
                            if constexpr (P::blk(b).isp) {
                                constexpr int ps = count_p<P>(b);
                                if constexpr (ASYNC) cl_st_async(paddr[ps][wi], lrot(cv, pswp[ps][wi]), s_vbar[ps]);
                                else cl_st(paddr[ps][wi], lrot(cv, pswp[ps][wi]));
                            } else {
                                idm[count_i<P>(b)][wi] = cv;
                            }
                        }
                    });
                });
            }
            // the v of all four CTAs are in this CTA's msg.  ASYNC: every word of msg is stored exactly once per
            // iteration, so the byte count is the event; no store of iteration i+1 can overtake a load of iteration i:
            // its sender first waits for the u this CTA derives from that load (and likewise for ubuf, through v).
            if constexpr (ASYNC) mbar_wait_cluster(vbar_sa, apar & 1u);
            else cl_sync();

            // ================= check phase (:391-405 and :422-447) =================
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                const int wv = tid + wi * NT;
                static_for<0, NROW>([&](auto ri) {
                    constexpr int r = decltype(ri)::value;
                    constexpr int DC = row_degree<P>(r);
                    uint32_t a[kMaxDegC], ck[kMaxDegC], mu[kMaxDegC];
                    uint32_t sx = 0;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            uint32_t cv;
                            if constexpr (P::blk(b).isp) cv = msg[count_p<P>(b) * S + wv];
                            else cv = idm[count_i<P>(b)][wi];
                            if constexpr (b == 2) {
                                bad[wi] = ~(hloc8[wi] ^ cv) & 0x80008000u;       // three biased sign bits XORed = NOT parity
                                cv &= 0x00ff00ffu;
                            }
                            const uint32_t old = cc[b][wi];
                            const uint32_t x = (cv ^ old) & (cv ^ (old + 0x00010001u));   // bit 7: sign flipped and old != 0
                            const uint32_t km = sign7_mask(x);
                            const uint32_t cor = (cv & ~km) | (0x007f007fu & km);         // killed -> v = 0
                            cc[b][wi] = cor;
                            ck[k] = cor;
                            a[k] = ch2u(__hsub2(cu2h(cor), cu2h(0x007f007fu)));           // -v as a signed fp16 lane
                            sx ^= cor;
                        }
                    });
                    min_excl_h<DC>(a, mu);
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            const uint32_t nm = sign7_mask(sx ^ ck[k]);
                            const uint32_t u = __vadd2(mu[k], nm) ^ nm;                    // +-mu, two's complement
                            if constexpr (P::blk(b).isp) {
                                const uint2 t = tab[(count_p<P>(b) * WPT + wi) * NT + tid];
                                if constexpr (ASYNC) cl_st_async(t.x, lrot(u, t.y), s_ubar[count_p<P>(b)]);
                                else cl_st(t.x, lrot(u, t.y));                             // push to the CTA that owns the variables
                            } else {
                                idm[count_i<P>(b)][wi] = u;
                            }
                        }
                    });
                });
            }
            // ---- exit test (:445-453): row 0 in the threads, rows 1..NROW-1 only if row 0 is clean everywhere ----
            uint32_t synd = 0;
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) synd |= bad[wi];
            hb_complete = false;
            bool row0_bad;
            if constexpr (ASYNC) {
                // the four flags travel like the messages and complete on the same barrier as the u of this phase; two
                // sets by iteration parity (a CTA is at most one barrier phase ahead of the slowest reader)
                const int local = __syncthreads_or(synd != 0);
                const uint32_t *fl = s_aflag[iter & 1u];
                if (tid < kCL) cl_st_async(cl_map(aflag_sa + ((iter & 1u) * kCL + rank) * 4u, (uint32_t)tid), (uint32_t)local, cl_map(ubar_sa, (uint32_t)tid));
                mbar_wait_cluster(ubar_sa, (apar >> 1) & 1u);
                apar ^= 3u;
                row0_bad = (fl[0] | fl[1] | fl[2] | fl[3]) != 0;
            } else {
                row0_bad = cluster_or(synd != 0);   // (its barrier also publishes the u pushed in this phase)
            }
            if (!row0_bad) {
                flush_pack();
                hb_complete = true;
                cl_sync();                      // every quarter's hard bits are in place
                // syndrome word `sw` (32 checks of row r, this CTA's quarter): identity blocks read the local words,
                // permutation blocks a 32-bit window of quarter (theta + rank) mod 4 of the block's column
                synd = 0;
                for (int sw = tid; sw < (NROW - 1) * QW; sw += NT) {
                    const int r = 1 + sw / QW, iq0 = (sw % QW) * 32;
                    uint32_t sy = 0;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if (P::blk(b).row == r) {
                            constexpr int col = P::blk(b).col;
                            if constexpr (P::blk(b).isp) {
                                const uint32_t qv = ((uint32_t)prm.theta[b] + rank) & 3u;
                                const int s = ((int)prm.phi[b][rank] + iq0) & (Q - 1);
                                const int w0 = s >> 5, w1 = (w0 + 1) & (QW - 1);
                                const uint32_t base = cl_map(hb_sa + (uint32_t)(col * QW) * 4u, qv);
                                sy ^= __funnelshift_r(cl_ld(base + (uint32_t)w0 * 4u), cl_ld(base + (uint32_t)w1 * 4u), s & 31);
                            } else {
                                sy ^= hb[col * QW + (iq0 >> 5)];
                            }
                        }
                    });
                    synd |= sy;
                }
                if (!cluster_or(synd != 0)) {
                    ok = true;
                    iters_run = iter;                                                      // :462
                    break;
                }
            }
        }
        if (!hb_complete) {              // decoding failed: the output is the hard decision of the last marginals (:466-473)
            flush_pack();
            __syncthreads();
        }

        // ---- output: this quarter of every column's hard decisions, MSB first (:455-461, :466-473) ----
        uint8_t *out = out_all + frame * (unsigned long long)(NV / 8);
        for (int i = tid; i < HBL; i += NT) {
            const int c = i / QW, w = i % QW;
            const uint32_t rev = __brev(hb[i]);
            uint8_t *o = out + ((size_t)c * M + rank * Q) / 8 + 4 * w;
            if ((reinterpret_cast<uintptr_t>(o) & 3u) == 0) {
                *reinterpret_cast<uint32_t *>(o) = __byte_perm(rev, 0, 0x0123);
            } else {
                o[0] = (uint8_t)(rev >> 24); o[1] = (uint8_t)(rev >> 16); o[2] = (uint8_t)(rev >> 8); o[3] = (uint8_t)rev;
            }
        }
        if (tid == 0 && rank == 0) {
            if (success) success[frame] = ok ? 1 : 0;
            if (iters_out) iters_out[frame] = iters_run;
        }
        cl_sync();                       // nobody still reads this CTA's hb / msg when the next frame re-initialises them
        cur ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Four messages per word.  The pair kernel above moves one 32-bit word per pair of edges across the cluster, 8
// significant bits in each 16-bit lane; its iteration is bound by the distributed-shared-memory stores
// (profiles/r02_cluster.md).  Here a thread owns the QUAD {t, t + Q/4, t + Q/2, t + 3Q/4} of its quarter -- the two
// word slots wi = 0, 1 of the pair layout with NT = Q/4 threads -- and a rotation by phi inside a quarter maps quads to
// quads: thread t's four bytes go to ONE word of thread (t -+ phi) mod Q/4, rotated by whole bytes.  So the two words
// of a block are packed to their low bytes (one PRMT), rotated (one SHF) and stored once; the receiver unpacks with one
// PRMT per word (zero fill for v, sign fill for u).  Half the stores, half the bytes, the same arithmetic.
// Block 2 (row 0's permutation block) keeps the pair form for v: its high bytes carry the marginals of the exit test.
template <int RATE, int M, int MINB, bool ASYNC>
__global__ void __cluster_dims__(kCL, 1, 1) __launch_bounds__(M / 16, MINB)
decode_ms_tm_cluster4_kernel(const TmParams prm, const int8_t *__restrict__ llrs_all, uint8_t *__restrict__ out_all,
                             unsigned long long batch, unsigned max_iters, uint8_t *__restrict__ success,
                             uint32_t *__restrict__ iters_out, const uint32_t one /* == 1, see above */) {
    typedef Proto<RATE> P;
    constexpr int NB = P::NB, NCOL = P::NCOL, NROW = P::NROW;
    constexpr int NP = count_p<P>(NB), NI = NB - NP;
    constexpr int WPT = 2;
    constexpr int Q = M / 4, S = Q / 2, NT = Q / 4;            // NT quads = S word slots per CTA
    constexpr int NV = NCOL * M, N = (NCOL - 1) * M;
    constexpr int QW = Q / 32, HBL = NCOL * QW;
    constexpr int CA = P::blk(0).col, CP = NCOL - 1;
    constexpr int NG = (NCOL + 1) / 2;
    static_assert(NT % 32 == 0, "whole warps");
    static_assert(P::blk(0).row == 0 && !P::blk(0).isp && P::blk(1).row == 0 && P::blk(1).col == CP && !P::blk(1).isp &&
                  P::blk(2).row == 0 && P::blk(2).col == CP && P::blk(2).isp && P::blk(3).row == 1 && count_p<P>(2) == 0,
                  "row 0 must be I(CA) + I(CP) + P(CP), and block 2 the first permutation block");

    extern __shared__ __align__(16) uint32_t smem_cl[];
    uint32_t *msg2 = smem_cl;                        // [S] v of block 2, pair form (check order)
    uint32_t *msg4 = msg2 + S;                       // [NP-1][NT] v of the other permutation blocks, quad form (check order)
    uint32_t *ubuf = msg4 + (NP - 1) * NT;           // [NP][NT] u, quad form (variable order)
    uint2 *tab = reinterpret_cast<uint2 *>(ubuf + NP * NT);   // [NP][NT] check side: {shared::cluster address in ubuf, left rotation}
    uint32_t *hb = reinterpret_cast<uint32_t *>(tab + NP * NT);
    constexpr unsigned FBL = (NCOL - 1) * Q;
    unsigned char *stage = reinterpret_cast<unsigned char *>(hb + ((HBL + 3) & ~3));   // [2][FBL]
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_flag[kCL];
    // ASYNC (see the pair kernel): [0] counts the bytes of v pushed into msg2 / msg4, [1] the bytes of u plus the four row-0 flags
    __shared__ __align__(8) uint64_t s_abar[2];
    __shared__ uint32_t s_vbar[NP], s_ubar[NP];
    __shared__ uint32_t s_aflag[2][kCL];
    constexpr uint32_t kVBytes = (uint32_t)(S + (NP - 1) * NT) * 4u, kUBytes = (uint32_t)NP * NT * 4u + 4u * kCL;

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t rank = cl_rank();
    const uint32_t msg2_sa = smem_addr(msg2), msg4_sa = smem_addr(msg4), ubuf_sa = smem_addr(ubuf), hb_sa = smem_addr(hb),
                   flag_sa = smem_addr(s_flag);
    const uint32_t vbar_sa = smem_addr(&s_abar[0]), ubar_sa = smem_addr(&s_abar[1]), aflag_sa = smem_addr(&s_aflag[0][0]);

    // per-thread constants of the variable side: where this thread's v of every permutation block goes
    uint32_t paddr[NP], prot[NP];                    // quad form: address, right rotation in bits
    uint32_t paddr2[WPT], pswp2[WPT];                // block 2, pair form
    static_for<0, NB>([&](auto bi) {
        constexpr int b = decltype(bi)::value;
        if constexpr (P::blk(b).isp) {
            constexpr int ps = count_p<P>(b);
            const int q = ((int)rank - (int)prm.theta[b]) & 3;               // the check quarter = the CTA that holds the message
            const int phi = prm.phi[b][q];                                   // variable j = (phi + check i) mod Q
            if constexpr (b == 2) {
#pragma unroll
                for (int wi = 0; wi < WPT; wi++) {
                    const int wv = tid + wi * NT;
                    const int borrow = wv < phi % S ? 1 : 0;
                    paddr2[wi] = cl_map(msg2_sa + (uint32_t)((wv - phi % S) & (S - 1)) * 4u, (uint32_t)q);
                    pswp2[wi] = ((phi / S ^ borrow) & 1) ? 16u : 0u;
                }
                paddr[ps] = 0; prot[ps] = 0;
            } else {
                const int borrow = tid < phi % NT ? 1 : 0;
                paddr[ps] = cl_map(msg4_sa + (uint32_t)((ps - 1) * NT + ((tid - phi % NT) & (NT - 1))) * 4u, (uint32_t)q);
                prot[ps] = 8u * (uint32_t)((phi / NT + borrow) & 3);          // variable byte k is the check's byte k - r
            }
            // the same block seen from the checks this thread owns (quarter `rank`): the variable quad they talk to
            const int qv2 = ((int)prm.theta[b] + (int)rank) & 3;
            const int phi2 = prm.phi[b][rank];
            const int t2 = tid + phi2 % NT;
            if (ASYNC && tid == 0) {
                s_vbar[ps] = cl_map(vbar_sa, (uint32_t)q);
                s_ubar[ps] = cl_map(ubar_sa, (uint32_t)qv2);
            }
            tab[ps * NT + tid] = make_uint2(cl_map(ubuf_sa + (uint32_t)(ps * NT + (t2 & (NT - 1))) * 4u, (uint32_t)qv2),
                                            8u * (uint32_t)((phi2 / NT + (t2 >= NT ? 1 : 0)) & 3));   // check byte k is the variable's byte k + r
        }
    });
    const uint32_t c255 = 0x00ff00ffu * one, c256 = one << 8;

    const unsigned long long n_clusters = cl_count();
    const bool use_bulk = (reinterpret_cast<uintptr_t>(llrs_all) & 15u) == 0;
    auto stage_frame = [&](unsigned long long frame, unsigned buf) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_bar[buf])), "r"(FBL) : "memory");
        const int8_t *src = llrs_all + frame * (unsigned long long)N + rank * Q;
#pragma unroll 1
        for (int c = 0; c < NCOL - 1; c++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_addr(stage + buf * FBL + c * Q)), "l"(src + (size_t)c * M), "r"((unsigned)Q),
                           "r"(smem_addr(&s_bar[buf])) : "memory");
    };
    unsigned long long frame = cl_id();
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        if (ASYNC) {
            mbar_init(&s_abar[0], 1);
            mbar_init(&s_abar[1], 1);
        }
        mbar_init_fence();
        if (use_bulk && frame < batch) stage_frame(frame, 0);
    }
    __syncthreads();
    if (ASYNC) cl_sync();
    unsigned cur = 0, bar_parity = 0;
    uint32_t apar = 0;

    for (; frame < batch; frame += n_clusters) {
        if (tid == 0 && use_bulk && frame + n_clusters < batch) stage_frame(frame + n_clusters, cur ^ 1);
        if (use_bulk) {
            mbar_wait(&s_bar[cur], (bar_parity >> cur) & 1u);
            bar_parity ^= 1u << cur;
        }
        const int8_t *llr_s = reinterpret_cast<const int8_t *>(stage + cur * FBL);
        const int8_t *llr_g = llrs_all + frame * (unsigned long long)N + rank * Q;

        // ---- per-frame state: everything zero, every call (:368, :374) ----
        uint32_t Lb[NCOL][WPT], idm[NI > 0 ? NI : 1][WPT], cc[NB][WPT];
#pragma unroll
        for (int wi = 0; wi < WPT; wi++) {
            const int wv = tid + wi * NT;
#pragma unroll
            for (int c = 0; c < NCOL; c++) {
                if (c < NCOL - 1) {
                    const int l0 = use_bulk ? llr_s[c * Q + wv] : llr_g[(size_t)c * M + wv];
                    const int l1 = use_bulk ? llr_s[c * Q + wv + S] : llr_g[(size_t)c * M + wv + S];
                    Lb[c][wi] = (uint32_t)(l0 + 128) | ((uint32_t)(l1 + 128) << 16);
                } else {
                    Lb[c][wi] = 0x00800080u;                                      // :383
                }
            }
#pragma unroll
            for (int i = 0; i < NI; i++) idm[i][wi] = 0;
#pragma unroll
            for (int b = 0; b < NB; b++) cc[b][wi] = 0;                  // v_old = +0 (fp16 lanes)
        }
#pragma unroll
        for (int p = 0; p < NP; p++) ubuf[p * NT + tid] = 0;            // u = 0 before the first iteration
        for (int i = tid; i < HBL; i += NT) hb[i] = 0;
        if (ASYNC) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cl_sync();

        unsigned iters_run = max_iters;
        bool ok = false, hb_complete = true;
        uint32_t gat[NG][WPT], hloc8[WPT], bad[WPT];
        auto cluster_or = [&](bool pred) {
            const int local = __syncthreads_or(pred);
            if (tid < kCL) cl_st(cl_map(flag_sa + rank * 4u, (uint32_t)tid), (uint32_t)local);
            cl_sync();
            return (s_flag[0] | s_flag[1] | s_flag[2] | s_flag[3]) != 0;
        };
        auto flush_pack = [&]() {
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) {
                const int w32 = (tid + wi * NT) >> 5;
                static_for<0, NCOL>([&](auto ci) {
                    constexpr int c = decltype(ci)::value;
                    const uint32_t g = gat[c / 2][wi] >> ((c & 1) * 8);
                    const unsigned b0 = __ballot_sync(0xFFFFFFFFu, (g & 0x00000080u) == 0);
                    const unsigned b1 = __ballot_sync(0xFFFFFFFFu, (g & 0x00800000u) == 0);
                    if (lane == 0) {
                        hb[c * QW + w32] = b0;
                        hb[c * QW + w32 + S / 32] = b1;
                    }
                });
            }
        };

        for (unsigned iter = 0; iter < max_iters; iter++) {
            if (ASYNC && tid == 0) {
                mbar_arm(vbar_sa, kVBytes);
                mbar_arm(ubar_sa, kUBytes);
            }
            // ================= variable phase (:382-411 and :421) =================
            static_for<0, NCOL>([&](auto ci) {
                constexpr int c = decltype(ci)::value;
                uint32_t hold[6];                                            // word 0's v of this column's blocks, until word 1's are there
#pragma unroll
                for (int wi = 0; wi < WPT; wi++) {
                    uint32_t va = Lb[c][wi];
                    uint32_t ub[6];
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            uint32_t u;
                            if constexpr (P::blk(b).isp) {
                                const uint32_t uq = ubuf[count_p<P>(b) * NT + tid];        // four signed bytes
                                u = wi == 0 ? sext_bytes02(uq) : sext_bytes13(uq);
                            } else {
                                u = idm[count_i<P>(b)][wi];
                            }
                            ub[k] = u;
                            va = __viaddmin_s16x2_relu(va, u, 0x00ff00ffu);
                        }
                    });
                    const uint32_t van = c255 * one - va;
                    if constexpr ((c & 1) == 0) gat[c / 2][wi] = va;
                    else gat[c / 2][wi] = __byte_perm(gat[c / 2][wi], va, 0x6240);
                    if constexpr (c == CA) hloc8[wi] = va;
                    if constexpr (c == CP) hloc8[wi] = (hloc8[wi] ^ va) * c256;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).col == c) {
                            constexpr int k = pos_in_col<P>(b);
                            uint32_t cv = __viaddmin_s16x2_relu(van, ub[k], 0x00fe00feu);
                            if constexpr (b == 2) {
                                cv = va * c256 + cv;
                                if constexpr (ASYNC) cl_st_async(paddr2[wi], lrot(cv, pswp2[wi]), s_vbar[0]);
                                else cl_st(paddr2[wi], lrot(cv, pswp2[wi]));
                            } else if constexpr (P::blk(b).isp) {
                                constexpr int ps = count_p<P>(b);
                                if (wi == 0) hold[k] = cv;
                                else {
                                    const uint32_t pk = __byte_perm(hold[k], cv, 0x6240);
                                    if constexpr (ASYNC) cl_st_async(paddr[ps], __funnelshift_r(pk, pk, prot[ps]), s_vbar[ps]);
                                    else cl_st(paddr[ps], __funnelshift_r(pk, pk, prot[ps]));
                                }
                            } else {
                                idm[count_i<P>(b)][wi] = cv * one + 0x64006400u;           // as fp16: 1024 + C
                            }
                        }
                    });
                }
            });
            if constexpr (ASYNC) mbar_wait_cluster(vbar_sa, apar & 1u);
            else cl_sync();

            // ================= check phase (:391-405 and :422-447) =================
            static_for<0, NROW>([&](auto ri) {
                constexpr int r = decltype(ri)::value;
                constexpr int DC = row_degree<P>(r);
                uint32_t uh[kMaxDegC];                                       // word 0's u of this row's blocks
#pragma unroll
                for (int wi = 0; wi < WPT; wi++) {
                    const int wv = tid + wi * NT;
                    uint32_t a[kMaxDegC], ck[kMaxDegC], mu[kMaxDegC];
                    uint32_t sx = 0;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            uint32_t cv;
                            if constexpr (b == 2) {
                                cv = msg2[wv];
                                bad[wi] = ~(hloc8[wi] ^ cv) & 0x80008000u;
                                cv = (cv & 0x00ff00ffu) | 0x64006400u;
                            } else if constexpr (P::blk(b).isp) {
                                const uint32_t pq = msg4[(count_p<P>(b) - 1) * NT + tid];   // four unsigned bytes
                                cv = wi == 0 ? __byte_perm(pq, 0x64646464u, 0x4240) : __byte_perm(pq, 0x64646464u, 0x4341);   // 0x6400 + C in both lanes
                            } else {
                                cv = idm[count_i<P>(b)][wi];
                            }
                            // the arithmetic of ARITH 10 (decode_ms_tm.cu): v = 1151 - (1024 + C) as an integer-valued fp16,
                            // keep = sat(v v_old + 1) is 0 exactly where the sign flipped and v_old != 0, v_cor = v keep + 0
                            const __half2 d = __hsub2(cu2h(0x647f647fu), cu2h(cv));
                            const __half2 kp = __hfma2_sat(d, cu2h(cc[b][wi]), cu2h(0x3c003c00u));
                            const uint32_t dc = ch2u(__hfma2(d, kp, cu2h(0u)));
                            cc[b][wi] = dc;
                            ck[k] = dc;
                            a[k] = dc;
                            sx ^= dc;                                                      // bit 15: product of signs
                        }
                    });
                    min_excl_h<DC>(a, mu);
                    sx = (sx & 0x80008000u) ^ 0x3c003c00u;                                 // +-1.0 in both lanes
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if constexpr (P::blk(b).row == r) {
                            constexpr int k = pos_in_row<P>(b);
                            // mu * +-1.0 + 1536 has the bit pattern 0x6600 +- mu: minus 0x6600 per lane = u in two's complement (written as
                            // an addition: __vsub2 makes ptxas build its constant in a register, wrongly in one instantiation)
                            const uint32_t pm = sx ^ (ck[k] & 0x80008000u);
                            const uint32_t u = __vadd2(ch2u(__hfma2(cu2h(mu[k]), cu2h(pm), cu2h(0x66006600u))), 0x9a009a00u);
                            if constexpr (P::blk(b).isp) {
                                if (wi == 0) {
                                    uh[k] = u;
                                } else {
                                    const uint2 t = tab[count_p<P>(b) * NT + tid];
                                    if constexpr (ASYNC) cl_st_async(t.x, lrot(__byte_perm(uh[k], u, 0x6240), t.y), s_ubar[count_p<P>(b)]);
                                    else cl_st(t.x, lrot(__byte_perm(uh[k], u, 0x6240), t.y));
                                }
                            } else {
                                idm[count_i<P>(b)][wi] = u;
                            }
                        }
                    });
                }
            });
            // ---- exit test (:445-453) ----
            uint32_t synd = 0;
#pragma unroll
            for (int wi = 0; wi < WPT; wi++) synd |= bad[wi];
            hb_complete = false;
            bool row0_bad;
            if constexpr (ASYNC) {
                const int local = __syncthreads_or(synd != 0);
                const uint32_t *fl = s_aflag[iter & 1u];
                if (tid < kCL) cl_st_async(cl_map(aflag_sa + ((iter & 1u) * kCL + rank) * 4u, (uint32_t)tid), (uint32_t)local, cl_map(ubar_sa, (uint32_t)tid));
                mbar_wait_cluster(ubar_sa, (apar >> 1) & 1u);
                apar ^= 3u;
                row0_bad = (fl[0] | fl[1] | fl[2] | fl[3]) != 0;
            } else {
                row0_bad = cluster_or(synd != 0);
            }
            if (!row0_bad) {
                flush_pack();
                hb_complete = true;
                cl_sync();
                synd = 0;
                for (int sw = tid; sw < (NROW - 1) * QW; sw += NT) {
                    const int r = 1 + sw / QW, iq0 = (sw % QW) * 32;
                    uint32_t sy = 0;
                    static_for<0, NB>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        if (P::blk(b).row == r) {
                            constexpr int col = P::blk(b).col;
                            if constexpr (P::blk(b).isp) {
                                const uint32_t qv = ((uint32_t)prm.theta[b] + rank) & 3u;
                                const int s = ((int)prm.phi[b][rank] + iq0) & (Q - 1);
                                const int w0 = s >> 5, w1 = (w0 + 1) & (QW - 1);
                                const uint32_t base = cl_map(hb_sa + (uint32_t)(col * QW) * 4u, qv);
                                sy ^= __funnelshift_r(cl_ld(base + (uint32_t)w0 * 4u), cl_ld(base + (uint32_t)w1 * 4u), s & 31);
                            } else {
                                sy ^= hb[col * QW + (iq0 >> 5)];
                            }
                        }
                    });
                    synd |= sy;
                }
                if (!cluster_or(synd != 0)) {
                    ok = true;
                    iters_run = iter;
                    break;
                }
            }
        }
        if (!hb_complete) {
            flush_pack();
            __syncthreads();
        }

        // ---- output (:455-461, :466-473) ----
        uint8_t *out = out_all + frame * (unsigned long long)(NV / 8);
        for (int i = tid; i < HBL; i += NT) {
            const int c = i / QW, w = i % QW;
            const uint32_t rev = __brev(hb[i]);
            uint8_t *o = out + ((size_t)c * M + rank * Q) / 8 + 4 * w;
            if ((reinterpret_cast<uintptr_t>(o) & 3u) == 0) {
                *reinterpret_cast<uint32_t *>(o) = __byte_perm(rev, 0, 0x0123);
            } else {
                o[0] = (uint8_t)(rev >> 24); o[1] = (uint8_t)(rev >> 16); o[2] = (uint8_t)(rev >> 8); o[3] = (uint8_t)rev;
            }
        }
        if (tid == 0 && rank == 0) {
            if (success) success[frame] = ok ? 1 : 0;
            if (iters_out) iters_out[frame] = iters_run;
        }
        cl_sync();
        cur ^= 1;
    }
}

template <int RATE, int M, int WPT, int MINB, bool ASYNC>
cudaError_t launch_cluster_v(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                           size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    typedef Proto<RATE> P;
    constexpr int NP = count_p<P>(P::NB);
    constexpr int Q = M / 4, S = Q / 2, NT = S / WPT;
    const TmParams prm = make_params<RATE>(c);
    // v buffer, u buffer, check-side address table, hard-bit words, two staging buffers
    const size_t smem = ((size_t)2 * NP * S + (size_t)2 * NP * S + (((size_t)P::NCOL * Q / 32 + 3) & ~(size_t)3)) * sizeof(uint32_t) +
                        2 * (size_t)(P::NCOL - 1) * Q;
    auto kern = decode_ms_tm_cluster_kernel<RATE, M, WPT, MINB, ASYNC>;
    static bool configured[kMaxDevices] = {};
    static int clusters_cached[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(kCL * ctx.sm_count), 1, 1);
        cfg.blockDim = dim3(NT, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) return e;
        clusters_cached[ctx.device] = n < 1 ? 1 : n;
        configured[ctx.device] = true;
    }
    unsigned long long clusters = (unsigned long long)clusters_cached[ctx.device];
    if (clusters > batch) clusters = batch;
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)(clusters * kCL), NT, smem, stream>>>(prm, static_cast<const int8_t *>(llrs), output,
                                                            (unsigned long long)batch, mi, success, iters, 1u);
    count_launch();
    return cudaGetLastError();
}

// LABRADOR_LDPC_CLUSTER_ASYNC=1: st.async + mbarrier instead of two cluster barriers per iteration (slower in the pair form:
// twice the stores, each paying its complete_tx; profiles/r02_cluster.md)
template <int RATE, int M, int WPT, int MINB>
cudaError_t launch_cluster(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                           size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    static const bool async = [] { const char *e = getenv("LABRADOR_LDPC_CLUSTER_ASYNC"); return e && atoi(e) != 0; }();
    return async ? launch_cluster_v<RATE, M, WPT, MINB, true>(ctx, c, llrs, output, batch, max_iters, success, iters, stream)
                 : launch_cluster_v<RATE, M, WPT, MINB, false>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
}


template <int RATE, int M, int MINB, bool ASYNC>
cudaError_t launch_cluster4_v(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                            size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    typedef Proto<RATE> P;
    constexpr int NP = count_p<P>(P::NB);
    constexpr int Q = M / 4, S = Q / 2, NT = Q / 4;
    const TmParams prm = make_params<RATE>(c);
    // v of block 2 (pairs), v of the other blocks and u (quads), check-side address table, hard-bit words, two staging buffers
    const size_t smem = ((size_t)S + (size_t)(NP - 1) * NT + (size_t)NP * NT + (size_t)2 * NP * NT +
                         (((size_t)P::NCOL * Q / 32 + 3) & ~(size_t)3)) * sizeof(uint32_t) + 2 * (size_t)(P::NCOL - 1) * Q;
    auto kern = decode_ms_tm_cluster4_kernel<RATE, M, MINB, ASYNC>;
    static bool configured[kMaxDevices] = {};
    static int clusters_cached[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(kCL * ctx.sm_count * MINB), 1, 1);
        cfg.blockDim = dim3(NT, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) return e;
        clusters_cached[ctx.device] = n < 1 ? 1 : n;
        configured[ctx.device] = true;
    }
    unsigned long long clusters = (unsigned long long)clusters_cached[ctx.device];
    if (clusters > batch) clusters = batch;
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)(clusters * kCL), NT, smem, stream>>>(prm, static_cast<const int8_t *>(llrs), output,
                                                            (unsigned long long)batch, mi, success, iters, 1u);
    count_launch();
    return cudaGetLastError();
}

template <int RATE, int M, int MINB>
cudaError_t launch_cluster4(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                            size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    // LABRADOR_LDPC_CLUSTER_ASYNC=0: two cluster barriers per iteration instead of st.async + mbarrier (A/B runs and tests)
    static const bool async = [] { const char *e = getenv("LABRADOR_LDPC_CLUSTER_ASYNC"); return !(e && atoi(e) == 0); }();
    return async ? launch_cluster4_v<RATE, M, MINB, true>(ctx, c, llrs, output, batch, max_iters, success, iters, stream)
                 : launch_cluster4_v<RATE, M, MINB, false>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
}

}  // namespace

// LABRADOR_LDPC_TM_CLUSTER=0 keeps the k = 16384 codes on the table-driven kernel (A/B runs and tests).
bool has_decode_ms_tm_cluster(int code) {
    static const bool off = [] { const char *e = getenv("LABRADOR_LDPC_TM_CLUSTER"); return e && atoi(e) == 0; }();
    return !off && code >= 9 && code <= 11;
}

bool launch_decode_ms_tm_cluster(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                                 uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err, const Front &front) {
    if (!has_decode_ms_tm_cluster(code) || front.kind != kFrontNone) return false;
    const CodeInfo &c = *code_info(code);
    // LABRADOR_LDPC_CLUSTER_QUAD=0: the two-messages-per-word version (A/B runs and tests)
    static const bool quad = [] { const char *e = getenv("LABRADOR_LDPC_CLUSTER_QUAD"); return !(e && atoi(e) == 0); }();
    static const int minb = [] { const char *e = getenv("LABRADOR_LDPC_CLUSTER_MINB"); return e ? atoi(e) : 2; }();   // CTAs per SM, codes 9 and 10
    if (quad) {
        switch (code) {
            case 9:
                if (!structure_matches<2>(c) || c.m != 2048) return false;
                *err = minb == 1 ? launch_cluster4<2, 2048, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream)
                                 : launch_cluster4<2, 2048, 2>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
                return true;
            case 10:
                if (!structure_matches<1>(c) || c.m != 4096) return false;
                *err = minb == 1 ? launch_cluster4<1, 4096, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream)
                                 : launch_cluster4<1, 4096, 2>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
                return true;
            case 11:
                if (!structure_matches<0>(c) || c.m != 8192) return false;
                *err = launch_cluster4<0, 8192, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
                return true;
            default:
                return false;
        }
    }
    switch (code) {
        case 9:
            if (!structure_matches<2>(c) || c.m != 2048) return false;
            *err = launch_cluster<2, 2048, 1, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
            return true;
        case 10:
            if (!structure_matches<1>(c) || c.m != 4096) return false;
            *err = launch_cluster<1, 4096, 1, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
            return true;
        case 11:
            if (!structure_matches<0>(c) || c.m != 8192) return false;
            *err = launch_cluster<0, 8192, 2, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream);
            return true;
        default:
            return false;
    }
}

}  // namespace ldpc
