// Min-sum decoder for the TC codes with i8 LLRs: TWO CODEWORDS PER REGISTER (one per 16-bit lane).
//
// Replaces LDPCCode::decode_ms::<i8> (reference src/decoder.rs:347-475) for TC128 / TC256 / TC512 on the
// batched path.  Structure as in decode_ms_tc.cu -- a group of G = min(M, 32) lanes owns element s (and s + 32
// for M = 64) of every prototype column and row, messages in a per-warp shared-memory slice in check order,
// __syncwarp() between the phases -- but every register and every shared-memory word carries the same
// element of two independent codewords in its two 16-bit halves, in the biased representation of the
// packed TM kernel (decode_ms_tm.cu: saturating_add and the variable->check message are one
// VIADDMNMX.S16x2.RELU each, |v| is one VABSDIFF4, minima are VIMNMX3.U16x2, the self-correction rule
// :422-426 is a bit test + PRMT lane mask).  The 16-bit lanes of an instruction never interact, so each
// half is exactly the scalar algorithm.  Each half runs its own stream of codewords: when the codeword
// of a half converges (or gives up) its output is written and the next frame is claimed into that half
// alone (masked re-initialisation) while the other half carries on -- iteration counts stay exact and a
// slow codeword costs its partner nothing.  Hard decisions of the two codewords share a byte (bits 0 / 1),
// so one XOR chain yields both parities of a check (:445-447).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "front.cuh"
#include "runtime.h"
#include "tc_common.cuh"

namespace ldpc {

namespace {

constexpr int kX2Warps = 4;

template <int M> __host__ __device__ constexpr int x2_msg_stride() { return M < 32 ? 32 * M + M : 32 * M; }

__device__ __forceinline__ __half2 x2_u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t x2_h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint32_t lane_mask_bit7(uint32_t x) {
    // bytes 0,1 <- sign of byte 0; bytes 2,3 <- sign of byte 2 (bit 7 of each 16-bit lane -> lane mask)
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xaa88;" : "=r"(r) : "r"(x));
    return r;
}

// u_k = min over the other seven edges of a check (= the reference's min1 / min2 selection, :391-395):
// prefixes and suffixes at pair boundaries, combined with the three-input minimum
__device__ __forceinline__ void min_excluding_self8(const uint32_t (&a)[8], uint32_t (&mu)[8]) {
    const uint32_t s3 = __vminu2(a[6], a[7]);
    const uint32_t s2 = __vimin3_u16x2(a[4], a[5], s3);
    const uint32_t s1 = __vimin3_u16x2(a[2], a[3], s2);
    mu[0] = __vminu2(a[1], s1);
    mu[1] = __vminu2(a[0], s1);
    const uint32_t p1 = __vminu2(a[0], a[1]);
    mu[2] = __vimin3_u16x2(p1, a[3], s2);
    mu[3] = __vimin3_u16x2(p1, a[2], s2);
    const uint32_t p2 = __vimin3_u16x2(p1, a[2], a[3]);
    mu[4] = __vimin3_u16x2(p2, a[5], s3);
    mu[5] = __vimin3_u16x2(p2, a[4], s3);
    const uint32_t p3 = __vimin3_u16x2(p2, a[4], a[5]);
    mu[6] = __vminu2(p3, a[7]);
    mu[7] = __vminu2(p3, a[6]);
}

// The same on fp16 lanes (as ARITH 6 of decode_ms_tm.cu): a lane holding the signed integer d (|d| < 1024) as the fp16
// subnormal d * 2^-24; |.| is a free operand modifier of HMNMX2 / VHMNMX, so d = C - 127 (one HADD2 on the FMA pipe) replaces
// the VABSDIFF4 on the ALU pipe, and the results are bit-identical to the integer |d|.
__device__ __forceinline__ uint32_t x2_hmin2a(uint32_t x, uint32_t y) {
    __half2 r = __hmin2(__habs2(*reinterpret_cast<__half2 *>(&x)), __habs2(*reinterpret_cast<__half2 *>(&y)));
    return *reinterpret_cast<uint32_t *>(&r);
}
__device__ __forceinline__ uint32_t x2_hmin3a(uint32_t x, uint32_t y, uint32_t z) {
    __half2 r = __hmin2(__hmin2(__habs2(*reinterpret_cast<__half2 *>(&x)), __habs2(*reinterpret_cast<__half2 *>(&y))),
                        __habs2(*reinterpret_cast<__half2 *>(&z)));
    return *reinterpret_cast<uint32_t *>(&r);
}
__device__ __forceinline__ void min_excluding_self8_h(const uint32_t (&a)[8], uint32_t (&mu)[8]) {
    const uint32_t s3 = x2_hmin2a(a[6], a[7]);
    const uint32_t s2 = x2_hmin3a(a[4], a[5], s3);
    const uint32_t s1 = x2_hmin3a(a[2], a[3], s2);
    mu[0] = x2_hmin2a(a[1], s1);
    mu[1] = x2_hmin2a(a[0], s1);
    const uint32_t p1 = x2_hmin2a(a[0], a[1]);
    mu[2] = x2_hmin3a(p1, a[3], s2);
    mu[3] = x2_hmin3a(p1, a[2], s2);
    const uint32_t p2 = x2_hmin3a(p1, a[2], a[3]);
    mu[4] = x2_hmin3a(p2, a[5], s3);
    mu[5] = x2_hmin3a(p2, a[4], s3);
    const uint32_t p3 = x2_hmin3a(p2, a[4], a[5]);
    mu[6] = x2_hmin2a(p3, a[7]);
    mu[7] = x2_hmin2a(p3, a[6]);
}

// MODE 0: integer lanes (VABSDIFF4 + VIMNMX); 1: |v| and the minima on fp16 lanes (ARITH 6 of decode_ms_tm.cu);
// 2: the self-correction rule and the sign of u on fp16 lanes as well (ARITH 10 of decode_ms_tm.cu: the variable side
// sends 0x6400 + C, v = 1151 - (1024 + C) is an integer-valued fp16, keep = sat(v v_old + 1), v_cor = v keep + 0,
// u = (mu * +-1.0 + 1536) - 0x6600 per lane) -- FMA-pipe instructions instead of LOP3 / PRMT on the ALU pipe;
// 3: MODE 2 with the hard decision of the sending variable in bit 15 of every message lane (|.| is a free operand
// modifier of the HADD2 that reads it), so the parity of a check (:445-447) is the XOR of the words it reads anyway: no
// hard-bit bytes in shared memory, no byte load per edge; the output bits are ballots of the marginals' sign bits
template <int M, int FRONT, int MODE>
__global__ void __launch_bounds__(32 * kX2Warps)
decode_ms_tc_i8x2_kernel(const typename FrontSrc<FRONT, int8_t>::type *__restrict__ llrs_all,
                         uint8_t *__restrict__ out_all, unsigned long long batch, unsigned max_iters,
                         uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out,
                         unsigned long long *__restrict__ counter, const float fscale, const float flimit,
                         const uint32_t one /* == 1: keeps an addition on the FMA pipe (IMAD) */) {
    constexpr bool HABS = MODE >= 1;
    constexpr int EPT = M > 32 ? M / 32 : 1;        // elements per lane
    constexpr int G = M / EPT;                       // lanes per codeword pair
    constexpr int CWW = 32 / G;                      // codeword pairs per warp
    constexpr int N = 8 * M;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int ESTRIDE = x2_msg_stride<M>();
    constexpr size_t kWarpBytes = 4u * CWW * ESTRIDE + (size_t)CWW * N;

    extern __shared__ __align__(16) unsigned char smem_x2[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cwl = lane / G, sl = lane % G;
    unsigned char *wbase = smem_x2 + (size_t)warp * ((kWarpBytes + 15) & ~(size_t)15);
    uint32_t *msg = reinterpret_cast<uint32_t *>(wbase) + (size_t)cwl * ESTRIDE;   // [32][M] words: half h = codeword h
    uint16_t *msg16 = reinterpret_cast<uint16_t *>(msg);
    uint8_t *hbv = wbase + 4u * CWW * ESTRIDE + (size_t)cwl * N;                   // [N] bytes: bit h = hard bit of codeword h
    const unsigned group_mask = (G == 32 ? kFull : ((1u << G) - 1u)) << (cwl * G);

    uint32_t Lv[8][EPT], vold[32][EPT];
    uint32_t hp[4][EPT];                             // MODE 3: bits 15 / 31 (even column) and 16 / 0 (odd column): hard decisions of the two codewords
#pragma unroll
    for (int ei = 0; ei < EPT; ei++) {
#pragma unroll
        for (int c = 0; c < 8; c++) Lv[c][ei] = 0x00800080u;
#pragma unroll
        for (int b = 0; b < 32; b++) { vold[b][ei] = MODE >= 2 ? 0u : 0x007f007fu; msg[b * M + sl + ei * G] = 0; }
    }
    bool have[2] = {false, false}, exhausted = false;
    unsigned long long frame[2] = {0, 0};
    unsigned iter[2] = {0, 0};

    for (;;) {
        // ---- every half without a codeword claims the next frame and (re-)initialises its 16-bit lanes ----
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const bool need = !have[h] && !exhausted;
            unsigned long long claimed = 0;
            if (sl == 0 && need) claimed = atomicAdd(counter, 1ull);
            claimed = __shfl_sync(kFull, claimed, cwl * G);
            if (need && claimed >= batch) exhausted = true;
            if (need && !exhausted) {
                frame[h] = claimed;
                have[h] = true;
                iter[h] = 0;
                const typename FrontSrc<FRONT, int8_t>::type *llr =
                    llrs_all + claimed * (unsigned long long)(FRONT == kFrontHard ? N / 8 : N);
                const uint32_t keep = h ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
                for (int ei = 0; ei < EPT; ei++) {
                    const int e = sl + ei * G;
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const int l = front_load<FRONT, int8_t>(llr, c * M + e, fscale, flimit);
                        Lv[c][ei] = (Lv[c][ei] & keep) | ((uint32_t)(l + 128) << (16 * h));
                    }
#pragma unroll
                    for (int b = 0; b < 32; b++) {                        // everything zero, every call (:368, :374)
                        vold[b][ei] = (vold[b][ei] & keep) | ((MODE >= 2 ? 0u : 0x7fu) << (16 * h));
                        msg16[(b * M + e) * 2 + h] = 0;
                    }
                }
            }
        }
        if (__all_sync(kFull, !have[0] && !have[1])) break;
        __syncwarp();

        // ---- variable phase (:382-411 and :421) ----
#pragma unroll
        for (int ei = 0; ei < EPT; ei++) {
            const int j = sl + ei * G;
            tc_static_for<0, 8>([&](auto ci) {
                constexpr int c = decltype(ci)::value;
                uint32_t va = Lv[c][ei];
                uint32_t ub[5];
                tc_static_for<0, 32>([&](auto bi) {
                    constexpr int b = decltype(bi)::value;
                    if constexpr (tc_blk(b).col == c) {
                        const int i = (j - tc_const_shift<M>(b)) & (M - 1);
                        const uint32_t u = msg[b * M + i];
                        ub[tc_pos_in_col(b)] = u;
                        va = __viaddmin_s16x2_relu(va, u, 0x00ff00ffu);             // saturating_add, ascending idx (:408)
                    }
                });
                // hard decision: va < 0  <=>  VA < 128  <=>  bit 7 clear
                const uint32_t nb = ~va;
                uint32_t magic = 0x64006400u;
                if constexpr (MODE == 3) {
                    const uint32_t hbit = (nb & 0x00800080u) * (one << 8);           // bits 15 / 31: the hard decisions
                    magic = hbit + 0x64006400u;                                      // as fp16: -(1024 + C) where the bit is 1
                    if constexpr ((c & 1) == 0) hp[c / 2][ei] = hbit;
                    else hp[c / 2][ei] += __funnelshift_l(hbit, hbit, 1);            // bits 16 / 0
                } else {
                    hbv[c * M + j] = (uint8_t)(((nb >> 7) & 1u) | ((nb >> 22) & 2u));
                }
                const uint32_t van = 0x00ff00ffu - va;
                tc_static_for<0, 32>([&](auto bi) {
                    constexpr int b = decltype(bi)::value;
                    if constexpr (tc_blk(b).col == c) {
                        const int i = (j - tc_const_shift<M>(b)) & (M - 1);
                        const uint32_t cv = __viaddmin_s16x2_relu(van, ub[tc_pos_in_col(b)], 0x00fe00feu);   // C = 127 - clamp(va - u)
                        msg[b * M + i] = MODE >= 2 ? cv * one + magic : cv;                                 // MODE 2, 3: as fp16, +-(1024 + C)
                    }
                });
            });
        }
        __syncwarp();

        // ---- check phase (:391-405 and :422-447) ----
        uint32_t par_any = 0;
#pragma unroll
        for (int ei = 0; ei < EPT; ei++) {
            const int i = sl + ei * G;
            tc_static_for<0, 4>([&](auto ri) {
                constexpr int r = decltype(ri)::value;
                uint32_t a[8], ck[8], mu[8];
                uint32_t sx = 0, par = 0;
                tc_static_for<0, 8>([&](auto ki) {
                    constexpr int k = decltype(ki)::value;
                    constexpr int b = r * 8 + k;
                    const uint32_t cv = msg[b * M + i];
                    if constexpr (MODE >= 2) {
                        const __half2 d = MODE == 3 ? __hsub2(x2_u2h(0x647f647fu), __habs2(x2_u2h(cv)))
                                                    : __hsub2(x2_u2h(0x647f647fu), x2_u2h(cv));       // v (0 -> +0)
                        const __half2 kp = __hfma2_sat(d, x2_u2h(vold[b][ei]), x2_u2h(0x3c003c00u));  // 0 where the sign flipped and v_old != 0
                        const uint32_t dc = x2_h2u(__hfma2(d, kp, x2_u2h(0u)));                       // killed -> +0
                        vold[b][ei] = dc;
                        ck[k] = dc;
                        a[k] = dc;
                        sx ^= dc;                                                   // bit 15: product of signs
                        if constexpr (MODE == 3) par ^= cv;                         // bits 15 / 31: parity of the hard decisions
                        else par ^= hbv[tc_blk(b).col * M + ((i + tc_const_shift<M>(b)) & (M - 1))];
                        return;
                    }
                    const uint32_t old = vold[b][ei];
                    const uint32_t x = (cv ^ old) & (cv ^ (old + 0x00010001u));     // bit 7: sign flipped and old != 0
                    const uint32_t km = lane_mask_bit7(x);
                    const uint32_t cor = (cv & ~km) | (0x007f007fu & km);           // killed -> v = 0
                    vold[b][ei] = cor;
                    ck[k] = cor;
                    if constexpr (HABS) {
                        const __half2 d = __hsub2(*reinterpret_cast<const __half2 *>(&cor), __half2(__ushort_as_half(0x007f), __ushort_as_half(0x007f)));
                        a[k] = *reinterpret_cast<const uint32_t *>(&d);             // -v as a signed fp16 lane
                    } else {
                        a[k] = __vabsdiffu4(cor, 0x007f007fu);                      // |v|
                    }
                    sx ^= cor;                                                      // bit 7: product of signs
                    par ^= hbv[tc_blk(b).col * M + ((i + tc_const_shift<M>(b)) & (M - 1))];
                });
                par_any |= par;
                if constexpr (HABS) min_excluding_self8_h(a, mu);
                else min_excluding_self8(a, mu);
                if constexpr (MODE >= 2) sx = (sx & 0x80008000u) ^ 0x3c003c00u;     // +-1.0 in both halves
                tc_static_for<0, 8>([&](auto ki) {
                    constexpr int k = decltype(ki)::value;
                    constexpr int b = r * 8 + k;
                    if constexpr (MODE >= 2) {
                        const uint32_t pm = sx ^ (ck[k] & 0x80008000u);             // +-1.0 with the sign of u (:398-405)
                        msg[b * M + i] = __vadd2(x2_h2u(__hfma2(x2_u2h(mu[k]), x2_u2h(pm), x2_u2h(0x66006600u))), 0x9a009a00u);
                    } else {
                        const uint32_t nm = lane_mask_bit7(sx ^ ck[k]);             // halves whose u is negative (:398-405)
                        msg[b * M + i] = __vadd2(mu[k], nm) ^ nm;                   // +-mu, two's complement per half
                    }
                });
            });
        }
        const unsigned bad0 = __ballot_sync(kFull, (par_any & (MODE == 3 ? 0x00008000u : 1u)) != 0) & group_mask;
        const unsigned bad1 = __ballot_sync(kFull, (par_any & (MODE == 3 ? 0x80000000u : 2u)) != 0) & group_mask;
        bool fin[2] = {false, false}, ok[2] = {false, false};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (have[h]) {
                if ((h ? bad1 : bad0) == 0) { fin[h] = true; ok[h] = true; }          // :453, :462 (iters = iter)
                else if (++iter[h] == max_iters) fin[h] = true;                       // :466-474 (iters = max_iters)
            }
        }
        // ---- output of the halves that are done: hard decisions MSB first (:455-461, :466-473) ----
        if constexpr (MODE == 3) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (!__any_sync(kFull, fin[h])) continue;                // warp-uniform: every lane takes part in the ballots
                uint8_t *out = out_all + frame[h] * (unsigned long long)(N / 8);
#pragma unroll
                for (int ei = 0; ei < EPT; ei++) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        // hard decision of element sl + ei * G of column c, codeword h (see the variable phase)
                        const uint32_t w = hp[c / 2][ei];
                        const bool hard = (c & 1) ? ((w >> (h ? 0 : 16)) & 1u) != 0 : ((w >> (h ? 31 : 15)) & 1u) != 0;
                        const unsigned bal = __ballot_sync(kFull, hard) >> (cwl * G);
                        if (fin[h] && sl < G / 8)                        // byte sl of this column: elements 8 sl .. 8 sl + 7, first element = MSB
                            out[(c * M + ei * G) / 8 + sl] = (uint8_t)(__brev((bal >> (8 * sl)) & 0xFFu) >> 24);
                    }
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (MODE != 3 && fin[h]) {
                have[h] = false;
                uint8_t *out = out_all + frame[h] * (unsigned long long)(N / 8);
                for (int o = sl; o < N / 8; o += G) {
                    unsigned byte = 0;
#pragma unroll
                    for (int bit = 0; bit < 8; bit++) byte |= (((unsigned)hbv[o * 8 + bit] >> h) & 1u) << (7 - bit);
                    out[o] = (uint8_t)byte;
                }
                if (sl == 0) {
                    if (success) success[frame[h]] = ok[h] ? 1 : 0;
                    if (iters_out) iters_out[frame[h]] = ok[h] ? iter[h] : max_iters;
                }
            }
            if (MODE == 3 && fin[h]) {
                have[h] = false;
                if (sl == 0) {
                    if (success) success[frame[h]] = ok[h] ? 1 : 0;
                    if (iters_out) iters_out[frame[h]] = ok[h] ? iter[h] : max_iters;
                }
            }
        }
        __syncwarp();
    }
}

template <int M, int FRONT, int MODE>
cudaError_t launch_x2h(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                      size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, const Front &front) {
    constexpr int EPT = M > 32 ? M / 32 : 1, G = M / EPT, CWW = 32 / G;
    const size_t warp_bytes = ((4u * CWW * x2_msg_stride<M>() + (size_t)CWW * 8 * M) + 15) & ~(size_t)15;
    const size_t smem = warp_bytes * kX2Warps;
    auto kern = decode_ms_tc_i8x2_kernel<M, FRONT, MODE>;
    static bool configured[kMaxDevices] = {};
    static int per_sm_cached[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kX2Warps, smem);
        if (e != cudaSuccess) return e;
        per_sm_cached[ctx.device] = per_sm < 1 ? 1 : per_sm;
        configured[ctx.device] = true;
    }
    const unsigned long long groups = (batch + 2 * CWW - 1) / (2 * CWW);
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm_cached[ctx.device];
    const unsigned long long need = (groups + kX2Warps - 1) / kX2Warps;
    if (grid > need) grid = need;
    WorkCounter wc(ctx, stream);
    if (wc.error() != cudaSuccess) return wc.error();
    unsigned long long *counter = wc.ptr();
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)grid, 32 * kX2Warps, smem, stream>>>(
        static_cast<const typename FrontSrc<FRONT, int8_t>::type *>(llrs), output, (unsigned long long)batch, mi,
        success, iters, counter, front.scale, front.limit, 1u);
    count_launch();
    return cudaGetLastError();
}

// LABRADOR_LDPC_TC_X2_HABS=0 keeps everything on integer lanes (VABSDIFF4 + VIMNMX), =1 moves |v| and the minima to fp16
// lanes, =2 the self-correction rule and the sign of u too (default for TC256 / TC512), =3 also the hard decisions inside
// the messages (default for TC128): A/B runs and tests.
template <int M, int FRONT>
cudaError_t launch_x2(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                      size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, const Front &front) {
    // default: 3 for TC128 (+1-5 %), 2 for TC256 / TC512 (MODE 3 is 0.5-1.6 % slower there: profiles/raw/r02ah_tc_x2_mode3_log.txt)
    static const int mode = [] { const char *e = getenv("LABRADOR_LDPC_TC_X2_HABS"); return e ? atoi(e) : (M == 16 ? 3 : 2); }();
    if (mode >= 3) return launch_x2h<M, FRONT, 3>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
    if (mode == 2) return launch_x2h<M, FRONT, 2>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
    if (mode == 1) return launch_x2h<M, FRONT, 1>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
    return launch_x2h<M, FRONT, 0>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
}

template <int M>
bool x2_dispatch(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                 uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err, const Front &front) {
    if (!tc_const_shifts_match<M>(c)) return false;
    switch (front.kind) {
        case kFrontNone: *err = launch_x2<M, kFrontNone>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front); return true;
        case kFrontSoftF32: *err = launch_x2<M, kFrontSoftF32>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front); return true;
        case kFrontHard: *err = launch_x2<M, kFrontHard>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front); return true;
        default: return false;
    }
}

}  // namespace

// i8 LLRs (native, quantised from f32 on load, or from hard bits) on a TC code, at least two codewords, max_iters > 0.
// LABRADOR_LDPC_TC_X2=0 keeps the one-codeword-per-lane-group kernel of decode_ms_tc.cu (A/B runs, tests).
bool tc_x2_enabled() {
    static const bool on = [] { const char *e = getenv("LABRADOR_LDPC_TC_X2"); return !e || atoi(e) != 0; }();
    return on;
}

bool launch_decode_ms_tc_x2(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                            uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err, const Front &front) {
    if (code < 0 || code > 2 || !tc_x2_enabled() || max_iters == 0 || batch < 2) return false;
    const CodeInfo &c = *code_info(code);
    if (!tc_structure_matches(c)) return false;
    switch (c.m) {
        case 16: return x2_dispatch<16>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 32: return x2_dispatch<32>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 64: return x2_dispatch<64>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, err, front);
        default: return false;
    }
}

}  // namespace ldpc
