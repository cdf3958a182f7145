// Batched hard <-> soft converters (HBM-bound streaming kernels).
//
// Replaces LDPCCode::hard_to_llrs<T> / llrs_to_hard<T>
// (reference src/decoder.rs:484-493, 498-509): bit 1 -> -one(), bit 0 -> +one(),
// MSB first within each byte; the inverse uses hard_bit (`< 0`).
#include <cuda_runtime.h>

#include "front.cuh"
#include "llr_arith.cuh"
#include "runtime.h"

namespace ldpc {
namespace {

constexpr int kThreads = 256;

// One thread per 16 output bytes (16 / sizeof(T) LLRs), so every store instruction of a warp covers 512
// contiguous bytes; unaligned caller buffers take the element-wise path.
template <class T>
__global__ void hard_to_llrs_kernel(const uint8_t *__restrict__ in, T *__restrict__ out,
                                    unsigned long long n_llrs, const bool vec_ok) {
    constexpr int ELEMS = 16 / sizeof(T);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_chunks = n_llrs / ELEMS;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_chunks; j += stride) {
        const unsigned long long e0 = j * ELEMS;
        unsigned bitsrc = in[e0 >> 3];
        if constexpr (ELEMS == 16) bitsrc = (bitsrc << 8) | in[(e0 >> 3) + 1];
        __align__(16) T vals[ELEMS];
#pragma unroll
        for (int b = 0; b < ELEMS; b++) {
            const int pos = ELEMS == 16 ? (15 - b) : (7 - (int)(e0 & 7) - b);       // MSB first within each byte
            vals[b] = (T)(((bitsrc >> pos) & 1) ? Arith<T>::neg(Arith<T>::one()) : Arith<T>::one());
        }
        T *dst = out + e0;
        if (vec_ok) {
            __stcs(reinterpret_cast<uint4 *>(dst), *reinterpret_cast<const uint4 *>(vals));      // written once, never re-read here
        } else {
#pragma unroll
            for (int b = 0; b < ELEMS; b++) dst[b] = vals[b];
        }
    }
}

// One thread per output byte: eight consecutive LLRs are fetched with 8- or 16-byte loads (when the
// caller's buffer is aligned), so a warp reads 32 * 8 * sizeof(T) contiguous bytes and writes 32.
template <class T>
__global__ void llrs_to_hard_kernel(const T *__restrict__ in, uint8_t *__restrict__ out,
                                    unsigned long long n_bytes, const bool vec_ok) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += stride) {
        __align__(16) T vals[8];
        const T *src = in + i * 8;
        if (vec_ok) {
            if constexpr (sizeof(T) == 1) {
                *reinterpret_cast<uint2 *>(vals) = __ldg(reinterpret_cast<const uint2 *>(src));
            } else {
#pragma unroll
                for (unsigned q = 0; q < sizeof(T) * 8 / 16; q++)
                    reinterpret_cast<uint4 *>(vals)[q] = __ldg(reinterpret_cast<const uint4 *>(src) + q);
            }
        } else {
#pragma unroll
            for (int b = 0; b < 8; b++) vals[b] = src[b];
        }
        unsigned byte = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) byte |= (Arith<T>::hard_bit(vals[b]) ? 1u : 0u) << (7 - b);
        out[i] = (uint8_t)byte;
    }
}

// i8 fast path: 16 LLRs (one 16-byte load) -> 2 output bytes; sign bits gathered with a multiply
// ("movemask": ((w >> 7) & 0x01010101) * 0x08040201 >> 24 = the four sign bits, first byte most significant).
__global__ void llrs_to_hard_i8x16_kernel(const uint4 *__restrict__ in, uint16_t *__restrict__ out,
                                          unsigned long long n_chunks) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks; i += stride) {
        const uint4 v = __ldg(in + i);
        auto nib = [](uint32_t w) { return (((w >> 7) & 0x01010101u) * 0x08040201u) >> 24; };
        const uint32_t b0 = (nib(v.x) << 4) | nib(v.y), b1 = (nib(v.z) << 4) | nib(v.w);
        out[i] = (uint16_t)(b0 | (b1 << 8));
    }
}

template <class T>
cudaError_t h2l(DeviceCtx &ctx, const uint8_t *in, void *out, unsigned long long n_bytes, cudaStream_t s) {
    if (n_bytes == 0) return cudaSuccess;
    const unsigned long long n_llrs = n_bytes * 8;
    const unsigned long long n_chunks = n_llrs / (16 / sizeof(T));
    unsigned long long blocks = (n_chunks + kThreads - 1) / kThreads;
    const unsigned long long cap = (unsigned long long)ctx.sm_count * 32;
    if (blocks > cap) blocks = cap;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    hard_to_llrs_kernel<T><<<(unsigned)blocks, kThreads, 0, s>>>(in, static_cast<T *>(out), n_llrs, vec_ok);
    count_launch();
    return cudaGetLastError();
}

template <class T>
cudaError_t l2h(DeviceCtx &ctx, const void *in, uint8_t *out, unsigned long long n_llrs, cudaStream_t s) {
    if (n_llrs == 0) return cudaSuccess;
    const unsigned long long n_bytes = n_llrs / 8;
    unsigned long long blocks = (n_bytes + kThreads - 1) / kThreads;
    const unsigned long long cap = (unsigned long long)ctx.sm_count * 32;
    if (blocks > cap) blocks = cap;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    if (sizeof(T) == 1 && vec_ok && (reinterpret_cast<uintptr_t>(out) & 1u) == 0) {
        const unsigned long long n_chunks = n_llrs / 16;        // n is a multiple of 128
        unsigned long long b2 = (n_chunks + kThreads - 1) / kThreads;
        if (b2 > cap) b2 = cap;
        llrs_to_hard_i8x16_kernel<<<(unsigned)b2, kThreads, 0, s>>>(static_cast<const uint4 *>(in),
                                                                     reinterpret_cast<uint16_t *>(out), n_chunks);
        count_launch();
        return cudaGetLastError();
    }
    llrs_to_hard_kernel<T><<<(unsigned)blocks, kThreads, 0, s>>>(static_cast<const T *>(in), out, n_bytes, vec_ok);
    count_launch();
    return cudaGetLastError();
}

// Stand-alone quantiser (the un-fused form of front.cuh's kFrontSoftF32): 16 soft values (four 16-byte
// loads) -> 16 LLRs per thread and iteration; scalar tail and unaligned buffers element-wise.
template <class T>
__global__ void quantise_kernel(const float *__restrict__ soft, T *__restrict__ out, unsigned long long count,
                                const float scale, const float limit, const bool vec_ok) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long n16 = vec_ok ? count / 16 : 0;
    for (unsigned long long j = tid; j < n16; j += stride) {
        __align__(16) T q[16];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const float4 f = reinterpret_cast<const float4 *>(soft)[j * 4 + v];
            q[4 * v + 0] = (T)quantise_soft(f.x, scale, limit);
            q[4 * v + 1] = (T)quantise_soft(f.y, scale, limit);
            q[4 * v + 2] = (T)quantise_soft(f.z, scale, limit);
            q[4 * v + 3] = (T)quantise_soft(f.w, scale, limit);
        }
#pragma unroll
        for (int v = 0; v < (int)sizeof(q) / 16; v++)
            reinterpret_cast<uint4 *>(out)[j * (sizeof(q) / 16) + v] = reinterpret_cast<const uint4 *>(q)[v];
    }
    for (unsigned long long i = n16 * 16 + tid; i < count; i += stride) out[i] = (T)quantise_soft(soft[i], scale, limit);
}

template <class T>
cudaError_t quantise(DeviceCtx &ctx, const float *soft, void *out, unsigned long long count, float scale, float limit,
                     cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    unsigned long long blocks = (count / 16 + kThreads - 1) / kThreads + 1;
    const unsigned long long cap = (unsigned long long)ctx.sm_count * 32;
    if (blocks > cap) blocks = cap;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(soft) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    quantise_kernel<T><<<(unsigned)blocks, kThreads, 0, s>>>(soft, static_cast<T *>(out), count, scale, limit, vec_ok);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_quantise(DeviceCtx &ctx, int code, int llr_type, const float *soft, void *llrs, size_t batch,
                            float scale, float limit, cudaStream_t stream) {
    const unsigned long long count = (unsigned long long)batch * ctx.codes[code].n;
    switch (llr_type) {
        case kI8: return quantise<int8_t>(ctx, soft, llrs, count, scale, limit, stream);
        case kI16: return quantise<int16_t>(ctx, soft, llrs, count, scale, limit, stream);
        default: return cudaErrorInvalidValue;
    }
}

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
