// Batched hard <-> soft converters (HBM-bound streaming kernels).
//
// Replaces LDPCCode::hard_to_llrs<T> / llrs_to_hard<T>
// (reference src/decoder.rs:484-493, 498-509): bit 1 -> -one(), bit 0 -> +one(),
// MSB first within each byte; the inverse uses hard_bit (`< 0`).
#include <cuda_runtime.h>

#include "llr_arith.cuh"
#include "runtime.h"

namespace ldpc {
namespace {

constexpr int kThreads = 256;

// One thread per input byte: eight LLRs written as one or more 16-byte stores.
template <class T>
__global__ void hard_to_llrs_kernel(const uint8_t *__restrict__ in, T *__restrict__ out,
                                    unsigned long long n_bytes) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += stride) {
        const unsigned byte = in[i];
        __align__(16) T vals[8];
#pragma unroll
        for (int b = 0; b < 8; b++)
            vals[b] = ((byte >> (7 - b)) & 1) ? Arith<T>::neg(Arith<T>::one()) : Arith<T>::one();
        T *dst = out + i * 8;
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (sizeof(T) * 8) % 16 == 0) {
            const uint4 *src4 = reinterpret_cast<const uint4 *>(vals);
#pragma unroll
            for (unsigned q = 0; q < sizeof(T) * 8 / 16; q++) reinterpret_cast<uint4 *>(dst)[q] = src4[q];
        } else if ((reinterpret_cast<uintptr_t>(dst) & 7u) == 0 && sizeof(T) == 1) {
            *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(vals);
        } else {
#pragma unroll
            for (int b = 0; b < 8; b++) dst[b] = vals[b];
        }
    }
}

// One thread per LLR; a warp ballot packs 32 hard bits, lanes 0..3 store one byte each.
template <class T>
__global__ void llrs_to_hard_kernel(const T *__restrict__ in, uint8_t *__restrict__ out,
                                    unsigned long long n_llrs) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // n_llrs is a multiple of 128, so whole warps are either in or out of range
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_llrs; i += stride) {
        const bool bit = Arith<T>::hard_bit(in[i]);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, bit);
        if (lane < 4) {
            const unsigned byte = __brev((m >> (8 * lane)) & 0xFFu) >> 24;
            out[((i - lane) >> 3) + lane] = (uint8_t)byte;   // warp base is a multiple of 32 LLRs = 4 bytes
        }
    }
}

template <class T>
cudaError_t h2l(DeviceCtx &ctx, const uint8_t *in, void *out, unsigned long long n_bytes, cudaStream_t s) {
    if (n_bytes == 0) return cudaSuccess;
    unsigned long long blocks = (n_bytes + kThreads - 1) / kThreads;
    const unsigned long long cap = (unsigned long long)ctx.sm_count * 16;
    if (blocks > cap) blocks = cap;
    hard_to_llrs_kernel<T><<<(unsigned)blocks, kThreads, 0, s>>>(in, static_cast<T *>(out), n_bytes);
    count_launch();
    return cudaGetLastError();
}

template <class T>
cudaError_t l2h(DeviceCtx &ctx, const void *in, uint8_t *out, unsigned long long n_llrs, cudaStream_t s) {
    if (n_llrs == 0) return cudaSuccess;
    unsigned long long blocks = (n_llrs + kThreads - 1) / kThreads;
    const unsigned long long cap = (unsigned long long)ctx.sm_count * 16;
    if (blocks > cap) blocks = cap;
    llrs_to_hard_kernel<T><<<(unsigned)blocks, kThreads, 0, s>>>(static_cast<const T *>(in), out, n_llrs);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_hard_to_llrs(DeviceCtx &ctx, int code, int llr_type, const uint8_t *input, void *llrs,
                                size_t batch, cudaStream_t stream) {
    const unsigned long long nb = (unsigned long long)batch * (ctx.codes[code].n / 8);
    switch (llr_type) {
        case kI8: return h2l<int8_t>(ctx, input, llrs, nb, stream);
        case kI16: return h2l<int16_t>(ctx, input, llrs, nb, stream);
        case kI32: return h2l<int32_t>(ctx, input, llrs, nb, stream);
        case kF32: return h2l<float>(ctx, input, llrs, nb, stream);
        case kF64: return h2l<double>(ctx, input, llrs, nb, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_llrs_to_hard(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output,
                                size_t batch, cudaStream_t stream) {
    const unsigned long long nl = (unsigned long long)batch * ctx.codes[code].n;
    switch (llr_type) {
        case kI8: return l2h<int8_t>(ctx, llrs, output, nl, stream);
        case kI16: return l2h<int16_t>(ctx, llrs, output, nl, stream);
        case kI32: return l2h<int32_t>(ctx, llrs, output, nl, stream);
        case kF32: return l2h<float>(ctx, llrs, output, nl, stream);
        case kF64: return l2h<double>(ctx, llrs, output, nl, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ldpc
