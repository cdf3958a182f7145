// Min-sum decoder for the TC codes (TC128 / TC256 / TC512), any LLR type: codewords per WARP.
//
// Replaces LDPCCode::decode_ms::<T> (reference src/decoder.rs:347-475) for the three telecommand
// codes.  Their parity-check matrix is a 4 x 8 array of M x M rotated identities (M = n/8 = 16, 32,
// 64; four cells hold the sum of two), 32 blocks, check degree 8, variable degree 5 or 3
// (reference src/codes/compact_parity_checks.rs:21-78, block order of src/codes/mod.rs:295-361).
//
// A group of G = min(M, 32) lanes decodes one codeword: lane s owns element s (and s+32 for M = 64)
// of every prototype column (variable side) and every prototype row (check side).  TC128 packs two
// codewords into one warp.  Messages live in a per-warp slice of shared memory in check order
// (the variable side reads/writes element (j - shift) mod M of the block; the rotations are compile-time constants, so
// equal ones share their index arithmetic), the two phases of an
// iteration are separated by __syncwarp() only -- no CTA barrier anywhere -- and every lane group
// claims its next codeword from an atomic counter the moment its current one converges or gives up,
// so a slow codeword stalls neither another warp nor the other group of its own warp.
// Arithmetic: f32 / f64 / i32 use the scalar DecodeFrom semantics of llr_arith.cuh, i8 / i16 the biased
// representation of biased_arith.cuh; saturating adds in ascending edge index per variable (:408);
// self-correction against the previous v kept in registers (:422-426); min over the other edges
// (= min1/min2 selection, :391-395).
#include <cuda_runtime.h>

#include <type_traits>

#include "biased_arith.cuh"
#include "front.cuh"
#include "llr_arith.cuh"
#include "runtime.h"
#include "tc_common.cuh"

namespace ldpc {

namespace {

constexpr int kWarpsPerCta = 4;

// message slots per codeword in a warp's shared-memory slice (32 M messages + bank-skew padding)
template <int M> __host__ __device__ constexpr int tc_msg_stride() { return M < 32 ? 32 * M + M : 32 * M; }

template <int M, class T, int FRONT = kFrontNone>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
decode_ms_tc_kernel(const typename FrontSrc<FRONT, T>::type *__restrict__ llrs_all,
                    uint8_t *__restrict__ out_all,
                    unsigned long long batch, unsigned max_iters, uint8_t *__restrict__ success,
                    uint32_t *__restrict__ iters_out, unsigned long long *__restrict__ counter,
                    const float fscale, const float flimit) {
    typedef Arith<T> A;
    typedef typename MsgStore<T>::type ST;          // shared-memory storage type of a message
    typedef typename A::C CT;                       // register (compute) type
    constexpr int EPT = M > 32 ? M / 32 : 1;        // elements per lane
    constexpr int G = M / EPT;                       // lanes per codeword
    constexpr int CWW = 32 / G;                      // codewords per warp
    constexpr int N = 8 * M, NC = 4 * M, E = 32 * M;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    // i8 / i16 run in the biased representation of biased_arith.cuh
    constexpr bool kBiased = std::is_same<T, int8_t>::value || std::is_same<T, int16_t>::value;
    constexpr int BITS = 8 * (int)sizeof(T);
    constexpr int kB = kBiased ? (1 << (BITS - 1)) : 0, kMaxV = kB - 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cwl = lane / G, sl = lane % G;
    // per-warp slice: messages [CWW][32][M] of T, then hard bits [CWW][N] bytes
    // the codewords of one warp start G message slots apart modulo 32, so that their lanes use disjoint banks
    constexpr int ESTRIDE = tc_msg_stride<M>();
    constexpr size_t kWarpBytes = sizeof(ST) * CWW * ESTRIDE + (size_t)CWW * N;
    unsigned char *wbase = smem_raw + (size_t)warp * ((kWarpBytes + 15) & ~(size_t)15);
    ST *msg = reinterpret_cast<ST *>(wbase) + (size_t)cwl * ESTRIDE;
    uint8_t *hbv = wbase + sizeof(ST) * CWW * ESTRIDE + (size_t)cwl * N;
    const unsigned group_mask = (G == 32 ? kFull : ((1u << G) - 1u)) << (cwl * G);

    // Every lane group runs its own stream of codewords: as soon as a group's codeword converges (or gives up)
    // the group writes it out and claims the next frame, while the other groups of the warp carry on with
    // theirs -- a slow codeword never idles its neighbour.  The warp executes one loop; `have` says whether
    // this group currently holds a codeword, `iter` is that codeword's iteration.
    CT Lv[8][EPT], vold[32][EPT];
    bool have = false, exhausted = false;
    unsigned long long frame = 0;
    unsigned iter = 0;
    for (;;) {
        const bool need = !have && !exhausted;
        unsigned long long claimed = 0;
        if constexpr (CWW == 1) {                         // one group per warp: `need` is warp-uniform
            if (need) {
                if (lane == 0) claimed = atomicAdd(counter, 1ull);
                claimed = __shfl_sync(kFull, claimed, 0);
            }
        } else {
            if (sl == 0 && need) claimed = atomicAdd(counter, 1ull);
            claimed = __shfl_sync(kFull, claimed, cwl * G);
        }
        if (need && claimed >= batch) exhausted = true;
        if (need && !exhausted) {
        frame = claimed;
        have = true;
        iter = 0;
        const typename FrontSrc<FRONT, T>::type *llr =
            llrs_all + frame * (unsigned long long)(FRONT == kFrontHard ? N / 8 : N);   // front.cuh
#pragma unroll
        for (int ei = 0; ei < EPT; ei++) {
            const int e = sl + ei * G;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                Lv[c][ei] = front_load<FRONT, T>(llr, c * M + e, fscale, flimit);
                if constexpr (kBiased) Lv[c][ei] += (CT)kB;
            }
#pragma unroll
            for (int b = 0; b < 32; b++) { vold[b][ei] = kBiased ? (CT)kMaxV : A::zero(); msg[b * M + e] = (ST)A::zero(); }
#pragma unroll
            for (int c = 0; c < 8; c++) hbv[c * M + e] = 0;
        }
        }
        if constexpr (CWW == 1) {
            if (!have) break;
        } else {
            if (__all_sync(kFull, !have)) break;         // every group has run out of frames
        }
        __syncwarp();

        bool finished = have && max_iters == 0;          // (false, 0) with an all-zero output
        bool ok = false;
        const bool run = have && !finished;
        {
            // ---- variable phase ----
            if (run) {
#pragma unroll
                for (int ei = 0; ei < EPT; ei++) {
                    const int j = sl + ei * G;
                    tc_static_for<0, 8>([&](auto ci) {
                        constexpr int c = decltype(ci)::value;
                        CT va = Lv[c][ei];
                        CT ub[5];
                        tc_static_for<0, 32>([&](auto bi) {
                            constexpr int b = decltype(bi)::value;
                            if constexpr (tc_blk(b).col == c) {
                                const int i = (j - tc_const_shift<M>(b)) & (M - 1);
                                const CT u = (CT)msg[b * M + i];
                                ub[tc_pos_in_col(b)] = u;
                                if constexpr (kBiased) va = (CT)__viaddmin_s32_relu((int)va, (int)u, 2 * kB - 1);
                                else va = A::sat_add(va, u);                             // :408
                            }
                        });
                        if constexpr (kBiased) hbv[c * M + j] = (int)va < kB ? 1 : 0;
                        else hbv[c * M + j] = A::hard_bit(va) ? 1 : 0;
                        [[maybe_unused]] const int van = kBiased ? (2 * kB - 1) - (int)va : 0;
                        tc_static_for<0, 32>([&](auto bi) {
                            constexpr int b = decltype(bi)::value;
                            if constexpr (tc_blk(b).col == c) {
                                const int i = (j - tc_const_shift<M>(b)) & (M - 1);
                                if constexpr (kBiased)
                                    msg[b * M + i] = (ST)__viaddmin_s32_relu(van, (int)ub[tc_pos_in_col(b)], 2 * kMaxV);
                                else
                                    msg[b * M + i] = (ST)A::sat_sub(va, ub[tc_pos_in_col(b)]);   // :421
                            }
                        });
                    });
                }
            }
            __syncwarp();
            // ---- check phase ----
            bool par_any = false;
            if (run) {
#pragma unroll
                for (int ei = 0; ei < EPT; ei++) {
                    const int i = sl + ei * G;
                    tc_static_for<0, 4>([&](auto ri) {
                        constexpr int r = decltype(ri)::value;
                        if constexpr (kBiased) {
                            uint32_t a[8], ck[8], mu[8];
                            uint32_t sx = 0;
                            int par = 0;
                            tc_static_for<0, 8>([&](auto ki) {
                                constexpr int k = decltype(ki)::value;
                                constexpr int b = r * 8 + k;
                                const uint32_t cv = (uint32_t)msg[b * M + i];
                                const uint32_t old = (uint32_t)vold[b][ei];
                                const uint32_t x = (cv ^ old) & (cv ^ (old + 1u));       // bit BITS-1: sign flipped and old != 0
                                const uint32_t km = sign_mask_of_byte<BITS / 8 - 1>(x);
                                const uint32_t cor = (cv & ~km) | ((uint32_t)kMaxV & km);   // killed -> v = 0  (:422-426)
                                vold[b][ei] = (CT)cor;
                                ck[k] = cor;
                                a[k] = __usad(cor, (uint32_t)kMaxV, 0u);                 // |v|
                                sx ^= cor;
                                par ^= hbv[tc_blk(b).col * M + ((i + tc_const_shift<M>(b)) & (M - 1))];   // :445-447
                            });
                            par_any |= par != 0;
                            min_excluding_self_u32<8, 8>(a, mu);                         // :391-395
                            tc_static_for<0, 8>([&](auto ki) {
                                constexpr int k = decltype(ki)::value;
                                constexpr int b = r * 8 + k;
                                const uint32_t nm = sign_mask_of_byte<BITS / 8 - 1>(sx ^ ck[k]);   // :398-405
                                msg[b * M + i] = (ST)((mu[k] + nm) ^ nm);
                            });
                        } else {
                        CT a[8], suf[8];
                        bool sg[8];
                        bool stot = false;
                        int par = 0;
                        tc_static_for<0, 8>([&](auto ki) {
                            constexpr int k = decltype(ki)::value;
                            constexpr int b = r * 8 + k;
                            const CT nv = (CT)msg[b * M + i];
                            const CT vo = vold[b][ei];
                            const bool keep = (A::hard_bit(nv) == A::hard_bit(vo)) || (vo == A::zero());
                            const CT v = keep ? nv : A::zero();                           // :422-426
                            vold[b][ei] = v;
                            a[k] = A::abs(v);
                            sg[k] = A::hard_bit(v);
                            stot ^= sg[k];
                            par ^= hbv[tc_blk(b).col * M + ((i + tc_const_shift<M>(b)) & (M - 1))];   // :445-447
                        });
                        par_any |= par != 0;
                        suf[7] = a[7];
#pragma unroll
                        for (int k = 6; k >= 1; k--) suf[k] = A::min(a[k], suf[k + 1]);
                        CT pre = a[0];
                        tc_static_for<0, 8>([&](auto ki) {
                            constexpr int k = decltype(ki)::value;
                            constexpr int b = r * 8 + k;
                            CT mu;
                            if constexpr (k == 0) mu = suf[1];
                            else if constexpr (k == 7) mu = pre;
                            else mu = A::min(pre, suf[k + 1]);
                            if constexpr (k > 0 && k < 7) pre = A::min(pre, a[k]);
                            if (stot != sg[k]) mu = A::neg(mu);                           // :398-405
                            msg[b * M + i] = (ST)mu;
                        });
                        }
                    });
                }
            }
            const unsigned bad = __ballot_sync(kFull, par_any);
            if (run) {
                if ((bad & group_mask) == 0) { finished = true; ok = true; }              // :453, :462 (iters = iter)
                else if (++iter == max_iters) finished = true;                            // :466-474 (iters = max_iters)
            }
            __syncwarp();
        }

        // ---- output: hard decisions MSB first (:455-461, :466-473); p = 0 for TC codes ----
        if (finished) {
            const unsigned iters_run = ok ? iter : (unsigned)max_iters;
            have = false;
            uint8_t *out = out_all + frame * (unsigned long long)(N / 8);
            for (int o = sl; o < N / 8; o += G) {
                unsigned byte = 0;
#pragma unroll
                for (int bit = 0; bit < 8; bit++) byte |= (unsigned)hbv[o * 8 + bit] << (7 - bit);
                out[o] = (uint8_t)byte;
            }
            if (sl == 0) {
                if (success) success[frame] = ok ? 1 : 0;
                if (iters_out) iters_out[frame] = iters_run;
            }
        }
        __syncwarp();
    }
}

template <int M, class T, int FRONT = kFrontNone>
cudaError_t launch_tc(DeviceCtx &ctx, const CodeInfo &c, const void *llrs, uint8_t *output, size_t batch,
                      size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                      const Front &front = Front()) {
    constexpr int EPT = M > 32 ? M / 32 : 1, G = M / EPT, CWW = 32 / G;
    const size_t warp_bytes = ((sizeof(typename MsgStore<T>::type) * CWW * tc_msg_stride<M>() + (size_t)CWW * 8 * M) + 15) & ~(size_t)15;
    const size_t smem = warp_bytes * kWarpsPerCta;
    auto kern = decode_ms_tc_kernel<M, T, FRONT>;
    static bool configured[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[ctx.device] = true;
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kWarpsPerCta, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    unsigned long long groups = (batch + CWW - 1) / CWW;
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm;
    const unsigned long long need = (groups + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need;
    WorkCounter wc(ctx, stream);
    if (wc.error() != cudaSuccess) return wc.error();
    unsigned long long *counter = wc.ptr();
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)grid, 32 * kWarpsPerCta, smem, stream>>>(
        static_cast<const typename FrontSrc<FRONT, T>::type *>(llrs), output, (unsigned long long)batch, mi, success,
        iters, counter, front.scale, front.limit);
    count_launch();
    return cudaGetLastError();
}

template <int M>
bool tc_dispatch(DeviceCtx &ctx, const CodeInfo &c, int llr_type, const void *llrs, uint8_t *output, size_t batch,
                 size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err,
                 const Front &front) {
    if (front.kind != kFrontNone) {      // fused front ends: (soft, i8 | i16) and (hard, i8)
        if (front.kind == kFrontSoftF32 && llr_type == kI8)
            *err = launch_tc<M, int8_t, kFrontSoftF32>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
        else if (front.kind == kFrontSoftF32 && llr_type == kI16)
            *err = launch_tc<M, int16_t, kFrontSoftF32>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
        else if (front.kind == kFrontHard && llr_type == kI8)
            *err = launch_tc<M, int8_t, kFrontHard>(ctx, c, llrs, output, batch, max_iters, success, iters, stream, front);
        else
            return false;
        return true;
    }
    switch (llr_type) {
        case kI8: *err = launch_tc<M, int8_t>(ctx, c, llrs, output, batch, max_iters, success, iters, stream); return true;
        case kI16: *err = launch_tc<M, int16_t>(ctx, c, llrs, output, batch, max_iters, success, iters, stream); return true;
        case kI32: *err = launch_tc<M, int32_t>(ctx, c, llrs, output, batch, max_iters, success, iters, stream); return true;
        case kF32: *err = launch_tc<M, float>(ctx, c, llrs, output, batch, max_iters, success, iters, stream); return true;
        case kF64: *err = launch_tc<M, double>(ctx, c, llrs, output, batch, max_iters, success, iters, stream); return true;
        default: return false;
    }
}

}  // namespace

bool launch_decode_ms_tc(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err,
                         const Front &front) {
    if (code < 0 || code > 2) return false;
    const CodeInfo &c = *code_info(code);
    if (!tc_structure_matches(c)) return false;
    // the kernels carry the rotations as compile-time constants (tc_common.cuh)
    if (!(c.m == 16 ? tc_const_shifts_match<16>(c) : c.m == 32 ? tc_const_shifts_match<32>(c) : tc_const_shifts_match<64>(c)))
        return false;
    switch (c.m) {
        case 16: return tc_dispatch<16>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 32: return tc_dispatch<32>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        case 64: return tc_dispatch<64>(ctx, c, llr_type, llrs, output, batch, max_iters, success, iters, stream, err, front);
        default: return false;
    }
}

bool has_decode_ms_tc(int code) { return code >= 0 && code <= 2; }

}  // namespace ldpc
