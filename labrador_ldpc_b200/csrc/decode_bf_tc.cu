// Bit-packed bit-flipping decoder for the TC codes (TC128 / TC256 / TC512): one codeword per THREAD.
//
// Replaces LDPCCode::decode_bf (reference src/decoder.rs:243-301) for the three telecommand codes (no
// punctured bits, so decode_erasures never runs, :252-257).  Their parity-check matrix is a 4 x 8 array
// of M x M rotated identities (M = 16, 32, 64; reference src/codes/compact_parity_checks.rs:21-78):
// check i of a block with shift s touches variable (i + s) mod M of the block's column
// (src/codes/mod.rs:305-311).  With every block column and block row held as one M-bit word
//   * the parities of a block row are the XOR of the row's eight column words, each rotated right by s;
//   * a block's contribution to the violated-check count of its column is the row's parity word rotated
//     left by s; the counts (<= 5) are summed as three bit planes by a carry-save adder tree;
//   * "flip every variable whose count equals the maximum" (:276-296) needs the set of counts that occur.
// A codeword is 16 / 32 / 64 bytes, so a whole decode fits in a thread's registers: no shared memory, no
// synchronisation, and consecutive threads read consecutive frames (16-byte loads).  Threads of a warp
// whose codeword has converged idle until the slowest one is done, and a codeword that never converges runs
// all max_iters iterations -- so large batches are decoded in two passes: the first stops after kFirstPassIters
// iterations and puts the few unfinished frames on a list, the second decodes those from scratch with the full
// iteration budget (same result: the decoder is deterministic and keeps no state between calls).
#include <cuda_runtime.h>

#include "bf_common.cuh"
#include "runtime.h"
#include "tc_common.cuh"

namespace ldpc {

namespace {

constexpr int kTcBfThreads = 128;
constexpr unsigned kFirstPassIters = 6;       // covers > 99 % of the frames that converge at all
constexpr size_t kTwoPassMinBatch = 1u << 18;   // below this a single pass is a few tens of microseconds anyway

template <int M> struct TcWord { typedef uint32_t type; };
template <> struct TcWord<64> { typedef uint64_t type; };

// bit i of the result = bit (i + s) mod M of x  (x has M significant bits; for M = 16 the caller passes
// x duplicated into both halves and ignores the upper half of the result)
template <int M> __device__ __forceinline__ typename TcWord<M>::type rot_right(typename TcWord<M>::type x, unsigned s) {
    if constexpr (M == 16) return x >> s;
    else if constexpr (M == 32) return __funnelshift_r(x, x, s);
    else return (x >> s) | (x << ((64u - s) & 63u));
}

// Decodes one frame with at most `iter_limit` iterations.  Returns false -- and writes nothing -- if the frame is
// still undecided after iter_limit < max_iters iterations (first pass); otherwise writes output / success / iters.
template <int M>
__device__ __forceinline__ bool decode_bf_tc_frame(const TcParams &prm, const uint8_t *__restrict__ in_all,
                                                   uint8_t *__restrict__ out_all, unsigned long long frame,
                                                   unsigned max_iters, unsigned iter_limit,
                                                   uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out) {
    typedef typename TcWord<M>::type W;
    constexpr int NWORDS = M / 4;                     // 32-bit words per codeword (n / 32)
    constexpr W kMask = M == 16 ? (W)0xFFFFu : ~(W)0;

    // ---- load: word k = variables 32k .. 32k+31, bit i = variable 32k + i (input is MSB first, :251) ----
    uint32_t r[NWORDS];
    const uint8_t *in = in_all + frame * (unsigned long long)(M);
    if ((reinterpret_cast<uintptr_t>(in_all) & 15u) == 0) {
#pragma unroll
        for (int q = 0; q < NWORDS / 4; q++) {
            const uint4 v = reinterpret_cast<const uint4 *>(in)[q];
            r[4 * q + 0] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < NWORDS; k++)
            r[k] = (uint32_t)in[4 * k] | ((uint32_t)in[4 * k + 1] << 8) | ((uint32_t)in[4 * k + 2] << 16) |
                   ((uint32_t)in[4 * k + 3] << 24);
    }
#pragma unroll
    for (int k = 0; k < NWORDS; k++) r[k] = __brev(__byte_perm(r[k], 0, 0x0123));

    W col[8];                                         // M = 16: kept duplicated in both halves
#pragma unroll
    for (int c = 0; c < 8; c++) {
        if constexpr (M == 16) col[c] = ((r[c >> 1] >> (16 * (c & 1))) & 0xFFFFu) * 0x10001u;
        else if constexpr (M == 32) col[c] = r[c];
        else col[c] = (W)r[2 * c] | ((W)r[2 * c + 1] << 32);
    }

    unsigned iters_run = max_iters;
    bool ok = false;
    for (unsigned iter = 0; iter < iter_limit; iter++) {
        // parity of every check (:269-273)
        W par[4] = {0, 0, 0, 0};
        tc_static_for<0, 32>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            par[tc_blk(b).row] ^= rot_right<M>(col[tc_blk(b).col], prm.shift[b]);
        });
        if constexpr (M == 16) {
#pragma unroll
            for (int rr = 0; rr < 4; rr++) par[rr] = (par[rr] & 0xFFFFu) * 0x10001u;
        }
        // violated-check count of every variable as bit planes (:276-286), and which counts occur
        W c0[8], c1[8], c2[8];
        W occurs[6] = {0, 0, 0, 0, 0, 0};
        tc_static_for<0, 8>([&](auto ci) {
            constexpr int c = decltype(ci)::value;
            constexpr int DEG = c < 4 ? 5 : 3;
            W x[6] = {0, 0, 0, 0, 0, 0};
            tc_static_for<0, 32>([&](auto bi) {
                constexpr int b = decltype(bi)::value;
                if constexpr (tc_blk(b).col == c)      // rotate left by s = rotate right by M - s
                    x[tc_pos_in_col(b)] = rot_right<M>(par[tc_blk(b).row], ((unsigned)M - prm.shift[b]) & (unsigned)(M - 1));
            });
            count_planes<DEG, W>(x, c0[c], c1[c], c2[c]);
#pragma unroll
            for (int v = 1; v <= DEG; v++)
                occurs[v] |= ((v & 1) ? c0[c] : ~c0[c]) & ((v & 2) ? c1[c] : ~c1[c]) & ((v & 4) ? c2[c] : ~c2[c]);
        });
        int max_viol = 0;
#pragma unroll
        for (int v = 1; v <= 5; v++) max_viol = (occurs[v] & kMask) ? v : max_viol;
        if (max_viol == 0) { ok = true; iters_run = iter; break; }              // :288-289
        // flip every variable whose count equals the maximum (:292-296)
        const W m0 = (max_viol & 1) ? ~(W)0 : (W)0, m1 = (max_viol & 2) ? ~(W)0 : (W)0, m2 = (max_viol & 4) ? ~(W)0 : (W)0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            W flip = ~((c0[c] ^ m0) | (c1[c] ^ m1) | (c2[c] ^ m2));
            if constexpr (M == 16) flip = (flip & 0xFFFFu) * 0x10001u;
            col[c] ^= flip;
        }
    }

    if (!ok && iter_limit < max_iters) return false;   // undecided: the second pass redoes this frame

    // ---- store: all n hard decisions, MSB first ----
#pragma unroll
    for (int c = 0; c < 8; c++) {
        if constexpr (M == 16) {
            if (c & 1) r[c >> 1] = (uint32_t)(col[c - 1] & 0xFFFFu) | ((uint32_t)col[c] << 16);
        } else if constexpr (M == 32) {
            r[c] = col[c];
        } else {
            r[2 * c] = (uint32_t)col[c];
            r[2 * c + 1] = (uint32_t)(col[c] >> 32);
        }
    }
#pragma unroll
    for (int k = 0; k < NWORDS; k++) r[k] = __byte_perm(__brev(r[k]), 0, 0x0123);
    uint8_t *out = out_all + frame * (unsigned long long)(M);
    if ((reinterpret_cast<uintptr_t>(out_all) & 15u) == 0) {
#pragma unroll
        for (int q = 0; q < NWORDS / 4; q++)
            reinterpret_cast<uint4 *>(out)[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < NWORDS; k++) {
            out[4 * k + 0] = (uint8_t)r[k];         out[4 * k + 1] = (uint8_t)(r[k] >> 8);
            out[4 * k + 2] = (uint8_t)(r[k] >> 16); out[4 * k + 3] = (uint8_t)(r[k] >> 24);
        }
    }
    if (success) success[frame] = ok ? 1 : 0;
    if (iters_out) iters_out[frame] = iters_run;
    return true;
}

// First (or only) pass: thread t decodes frame t; frames left undecided go on `list` (list[0] = count).
template <int M>
__global__ void __launch_bounds__(kTcBfThreads)
decode_bf_tc_kernel(const TcParams prm, const uint8_t *__restrict__ in_all, uint8_t *__restrict__ out_all,
                    unsigned long long batch, unsigned max_iters, unsigned iter_limit, uint8_t *__restrict__ success,
                    uint32_t *__restrict__ iters_out, unsigned *__restrict__ list) {
    const unsigned long long frame = (unsigned long long)blockIdx.x * kTcBfThreads + threadIdx.x;
    if (frame >= batch) return;
    if (!decode_bf_tc_frame<M>(prm, in_all, out_all, frame, max_iters, iter_limit, success, iters_out))
        list[1 + atomicAdd(&list[0], 1u)] = (unsigned)frame;
}

// Second pass: the listed frames, full iteration budget.
template <int M>
__global__ void __launch_bounds__(kTcBfThreads)
decode_bf_tc_retry_kernel(const TcParams prm, const uint8_t *__restrict__ in_all, uint8_t *__restrict__ out_all,
                          unsigned max_iters, uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out,
                          const unsigned *__restrict__ list) {
    const unsigned count = list[0];
    for (unsigned i = blockIdx.x * kTcBfThreads + threadIdx.x; i < count; i += gridDim.x * kTcBfThreads)
        decode_bf_tc_frame<M>(prm, in_all, out_all, list[1 + i], max_iters, max_iters, success, iters_out);
}

template <int M>
cudaError_t launch_bf_tc(DeviceCtx &ctx, const CodeInfo &c, const uint8_t *input, uint8_t *output, size_t batch, size_t max_iters,
                         uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    TcParams prm{};
    for (int b = 0; b < 32; b++) prm.shift[b] = (uint8_t)c.blocks[b].shift;
    const unsigned long long grid = (batch + kTcBfThreads - 1) / kTcBfThreads;
    if (grid > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    if (batch < kTwoPassMinBatch || batch > 0xFFFFFFF0ull || mi <= kFirstPassIters) {
        decode_bf_tc_kernel<M><<<(unsigned)grid, kTcBfThreads, 0, stream>>>(prm, input, output, (unsigned long long)batch,
                                                                             mi, mi, success, iters, nullptr);
        count_launch();
        return cudaGetLastError();
    }
    // two passes.  The list of undecided frames is a grow-only buffer of the device context (launchers run under the
    // context mutex); an event makes a call on another stream wait for the previous user of the list.
    const size_t need = (batch + 1) * sizeof(unsigned);
    cudaError_t e = cudaSuccess;
    if (!ctx.retry_done) {
        e = cudaEventCreateWithFlags(&ctx.retry_done, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    } else {
        e = cudaStreamWaitEvent(stream, ctx.retry_done, 0);
        if (e != cudaSuccess) return e;
    }
    if (ctx.retry_list_bytes < need) {
        if (ctx.retry_list) {
            cudaEventSynchronize(ctx.retry_done);
            cudaFree(ctx.retry_list);
            ctx.retry_list = nullptr;
            ctx.retry_list_bytes = 0;
        }
        e = cudaMalloc(reinterpret_cast<void **>(&ctx.retry_list), need);
        if (e != cudaSuccess) return e;
        ctx.retry_list_bytes = need;
    }
    unsigned *list = ctx.retry_list;
    e = cudaMemsetAsync(list, 0, sizeof(unsigned), stream);
    if (e == cudaSuccess) {
        decode_bf_tc_kernel<M><<<(unsigned)grid, kTcBfThreads, 0, stream>>>(prm, input, output, (unsigned long long)batch,
                                                                             mi, kFirstPassIters, success, iters, list);
        unsigned long long grid2 = (unsigned long long)ctx.sm_count * 8;
        if (grid2 > grid) grid2 = grid;
        decode_bf_tc_retry_kernel<M><<<(unsigned)grid2, kTcBfThreads, 0, stream>>>(prm, input, output, mi, success, iters, list);
        count_launch(2);
        e = cudaGetLastError();
    }
    const cudaError_t e2 = cudaEventRecord(ctx.retry_done, stream);
    return e != cudaSuccess ? e : e2;
}

}  // namespace

// Returns true (and launches) for the TC codes.
bool launch_decode_bf_tc(DeviceCtx &ctx, int code, const uint8_t *input, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err) {
    if (code < 0 || code > 2) return false;
    const CodeInfo &c = *code_info(code);
    if (!tc_structure_matches(c)) return false;
    switch (c.m) {
        case 16: *err = launch_bf_tc<16>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 32: *err = launch_bf_tc<32>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        case 64: *err = launch_bf_tc<64>(ctx, c, input, output, batch, max_iters, success, iters, stream); return true;
        default: return false;
    }
}

}  // namespace ldpc
