// Batched systematic encoders that work from the compact generator: the bit-packed GF(2) XOR kernel
// (encode_kernel) and the TC-code table kernel (encode_tc_lut_kernel).  The TM codes are encoded through the
// sparse parity-check matrix instead (encode_tm.cu); launch_encode() below picks the kernel.
//
// Replaces EncodeInto::encode_parity for u8/u32/u64 and LDPCCode::encode /
// copy_encode (reference src/encoder.rs:41-82, 107-160, 189-252, 292-315).
//
// The reference XORs generator row `crow` into the parity for every set data
// bit crow*b + o and rotates each b-bit parity block left once per offset, so
// the row contributed by data bit (crow, o) is the compact row gc[crow] with
// every b-bit block rotated RIGHT by o.
//   * encode_kernel: one thread owns one (or WPT) 32-bit parity word(s) (MSB = lowest bit index, the
//     reference's byte order) and, for every set data bit, XORs in the matching 32-bit window of the rotated
//     row, taken with one funnel shift from two adjacent words of the row's block.  The compact generator
//     (<= 4 KB) and the frame's data words sit in shared memory.  Small TC batches; A/B reference for the rest.
//   * encode_tc_lut_kernel (TC128) / encode_tc_rot_kernel (TC256, TC512): one codeword per thread over a table of
//     per-nibble / per-byte parity contributions (see the comments at the kernels).
// No tensor cores: this is GF(2), not a real-valued contraction.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "runtime.h"

namespace ldpc {
namespace {

__device__ __forceinline__ uint32_t load_be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

// grid.x CTAs; each CTA handles `fpc` frames at a time.  A frame is encoded by TPC = W / WPT threads;
// thread t owns the WPT parity words t, t + TPC, t + 2 TPC, ... -- TPC is a multiple of the words per
// circulant block, so all of a thread's words sit at the same position inside their blocks and share the
// data-bit scan (clz / clear) and the generator-word indices; only the funnel shift + XOR are per word.
template <int WPT>
__global__ void encode_kernel(const DeviceCode code, const uint8_t *__restrict__ data_all,
                              uint8_t *__restrict__ cw_all, unsigned long long batch, int fpc) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int k = code.k, n = code.n, b = code.b;
    const int r = n - k;
    const int W = r / 32;              // parity words per frame
    const int KW = k / 32;             // data words per frame
    const int crows = k / b;
    const int TPC = W / WPT;           // threads per frame
    uint32_t *gen_s = reinterpret_cast<uint32_t *>(smem);     // [crows][W]
    uint32_t *dw_s = gen_s + crows * W;                        // [fpc][KW]
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int i = tid; i < crows * W; i += nt) gen_s[i] = code.gen32[i];

    const int slot = tid / TPC, t = tid % TPC;
    const bool active = slot < fpc;
    const unsigned long long n_groups = (batch + fpc - 1) / fpc;
    const bool aligned = ((reinterpret_cast<uintptr_t>(cw_all) | (data_all ? reinterpret_cast<uintptr_t>(data_all) : 0)) & 3u) == 0;

    for (unsigned long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const unsigned long long f0 = g * (unsigned long long)fpc;
        __syncthreads();   // gen_s ready / previous group's dw_s no longer read
        // stage data words (MSB-first) and, for copy_encode, copy data into the codeword
        for (int i = tid; i < fpc * KW; i += nt) {
            const unsigned long long f = f0 + i / KW;
            if (f < batch) {
                const int wi = i % KW;
                const uint8_t *src = data_all ? data_all + f * (unsigned long long)(k / 8) + 4 * wi
                                              : cw_all + f * (unsigned long long)(n / 8) + 4 * wi;
                uint8_t *dst = cw_all + f * (unsigned long long)(n / 8) + 4 * wi;
                if (aligned) {
                    const uint32_t raw = *reinterpret_cast<const uint32_t *>(src);
                    dw_s[i] = __byte_perm(raw, 0, 0x0123);
                    if (data_all) *reinterpret_cast<uint32_t *>(dst) = raw;
                } else {
                    const uint32_t d = load_be32(src);
                    dw_s[i] = d;
                    if (data_all) {
                        dst[0] = (uint8_t)(d >> 24); dst[1] = (uint8_t)(d >> 16);
                        dst[2] = (uint8_t)(d >> 8);  dst[3] = (uint8_t)d;
                    }
                }
            }
        }
        __syncthreads();
        const unsigned long long f = f0 + slot;
        if (active && f < batch) {
            const uint32_t *dw = dw_s + slot * KW;
            uint32_t acc[WPT];
#pragma unroll
            for (int j = 0; j < WPT; j++) acc[j] = 0;
            if (b >= 32) {
                const int nb = b / 32;                 // words per circulant block
                const int wi = t % nb;                 // word inside the block (same for all owned words)
                for (int crow = 0; crow < crows; crow++) {
                    const uint32_t *grow = gen_s + crow * W + (t - wi);     // first owned block of this row
                    for (int q = 0; q < nb; q++) {
                        uint32_t D = dw[crow * nb + q];          // data bits o = 32q .. 32q+31 of this row
                        if (D == 0) continue;
                        int ia = wi - q; if (ia < 0) ia += nb;
                        int ib = ia - 1; if (ib < 0) ib += nb;
                        uint32_t X[WPT], Y[WPT];
#pragma unroll
                        for (int j = 0; j < WPT; j++) { Y[j] = grow[j * TPC + ia]; X[j] = grow[j * TPC + ib]; }
                        while (D) {
                            const int o2 = __clz(D);
                            D &= ~(0x80000000u >> o2);
#pragma unroll
                            for (int j = 0; j < WPT; j++)
                                acc[j] ^= __funnelshift_r(Y[j], X[j], o2);   // window of the row rotated right by 32q+o2
                        }
                    }
                }
            } else {
                // b == 16 (TC128): two 16-bit circulant blocks per parity word, 16 data bits per row (WPT == 1)
                for (int crow = 0; crow < crows; crow++) {
                    const uint32_t row = gen_s[crow * W + t];
                    const uint32_t hi = row >> 16, lo = row & 0xFFFFu;
                    const uint32_t dword = dw[crow / 2];
                    uint32_t D = (crow & 1) ? (dword & 0xFFFFu) : (dword >> 16);   // MSB-first 16 bits
                    while (D) {
                        const int o = __clz(D) - 16;
                        D &= ~(0x8000u >> o);
                        const uint32_t rh = ((hi >> o) | (hi << (16 - o))) & 0xFFFFu;
                        const uint32_t rl = ((lo >> o) | (lo << (16 - o))) & 0xFFFFu;
                        acc[0] ^= (rh << 16) | rl;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < WPT; j++) {
                const int w = t + j * TPC;
                uint8_t *dst = cw_all + f * (unsigned long long)(n / 8) + k / 8 + 4 * w;
                if (aligned) {
                    *reinterpret_cast<uint32_t *>(dst) = __byte_perm(acc[j], 0, 0x0123);
                } else {
                    dst[0] = (uint8_t)(acc[j] >> 24); dst[1] = (uint8_t)(acc[j] >> 16);
                    dst[2] = (uint8_t)(acc[j] >> 8);  dst[3] = (uint8_t)acc[j];
                }
            }
        }
    }
}


// ---- TC codes, large batches: one codeword per thread over a lookup table ---------------------------------
// A TC codeword is 16 / 32 / 64 bytes.  The parity contribution of every value of every data byte (TC128, TC256:
// 16 / 64 KB) or nibble (TC512: 32 KB) is precomputed (code_tables.h: tc_encoder_lut) in memory byte order, so a
// thread loads its data words, XORs k/8 (k/4) table rows selected by the data bytes and stores the codeword:
// no bit scans, no rotations, no byte swaps, no synchronisation.  KW = k/32 = (n-k)/32 words; GB = bits per group.
template <int KW, int GB>
__global__ void __launch_bounds__(512, 2)
encode_tc_lut_kernel(const uint32_t *__restrict__ lut_g, const uint8_t *__restrict__ data_all,
                     uint8_t *__restrict__ cw_all, unsigned long long batch, const uint32_t row_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NV = 1 << GB, GROUPS = KW * 32 / GB;
    constexpr int LUTW = GROUPS * NV * KW;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(lut_g);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < LUTW / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t lut_sa = (uint32_t)__cvta_generic_to_shared(smem);
    typedef typename std::conditional<KW == 2, uint2, uint4>::type Vec;      // widest vector that divides a frame half
    constexpr int VW = sizeof(Vec) / 4, NVEC = KW / VW;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(cw_all) | reinterpret_cast<uintptr_t>(data_all)) & (sizeof(Vec) - 1)) == 0;
    const unsigned long long in_stride = data_all ? KW * 4ull : KW * 8ull;
    const uint8_t *in_base = data_all ? data_all : cw_all;

    for (unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; f < batch;
         f += (unsigned long long)gridDim.x * blockDim.x) {
        const uint8_t *in = in_base + f * in_stride;
        uint8_t *cw = cw_all + f * (KW * 8ull);
        uint32_t d[KW], p[KW];
        if (vec_ok) {
#pragma unroll
            for (int i = 0; i < NVEC; i++) {
                const Vec v = reinterpret_cast<const Vec *>(in)[i];
                memcpy(&d[i * VW], &v, sizeof(Vec));
            }
        } else {
#pragma unroll
            for (int w = 0; w < KW; w++)
                d[w] = (uint32_t)in[4 * w] | ((uint32_t)in[4 * w + 1] << 8) | ((uint32_t)in[4 * w + 2] << 16) | ((uint32_t)in[4 * w + 3] << 24);
        }
#pragma unroll
        for (int w = 0; w < KW; w++) p[w] = 0;
#pragma unroll
        for (int w = 0; w < KW; w++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t byte = __byte_perm(d[w], 0, 0x4440 + b);
                if constexpr (GB == 8) {
                    const uint32_t a = lut_sa + (uint32_t)((4 * w + b) * NV * KW * 4) + byte * row_bytes;
#pragma unroll
                    for (int i = 0; i < NVEC; i++) {
                        Vec v;
                        if constexpr (KW == 2) asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
                        else asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a + 16 * i));
                        uint32_t t[VW];
                        memcpy(t, &v, sizeof(Vec));
#pragma unroll
                        for (int j = 0; j < VW; j++) p[i * VW + j] ^= t[j];
                    }
                } else {
                    // high nibble = group 2j, low nibble = group 2j + 1 of byte j = 4w + b
                    const uint32_t a_hi = lut_sa + (uint32_t)(((4 * w + b) * 2) * NV * KW * 4) + (byte >> 4) * row_bytes;
                    const uint32_t a_lo = lut_sa + (uint32_t)(((4 * w + b) * 2 + 1) * NV * KW * 4) + (byte & 15u) * row_bytes;
                    if constexpr (KW == 2) {
                        uint2 x, y;
                        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x.x), "=r"(x.y) : "r"(a_hi));
                        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(y.x), "=r"(y.y) : "r"(a_lo));
                        p[0] ^= x.x ^ y.x; p[1] ^= x.y ^ y.y;
                    } else {
#pragma unroll
                        for (int i = 0; i < NVEC; i++) {
                            uint4 x, y;
                            asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(a_hi + 16 * i));
                            asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(y.x), "=r"(y.y), "=r"(y.z), "=r"(y.w) : "r"(a_lo + 16 * i));
                            p[i * 4 + 0] ^= x.x ^ y.x; p[i * 4 + 1] ^= x.y ^ y.y; p[i * 4 + 2] ^= x.z ^ y.z; p[i * 4 + 3] ^= x.w ^ y.w;
                        }
                    }
                }
            }
        }
        if (vec_ok) {
#pragma unroll
            for (int i = 0; i < NVEC; i++) {
                Vec v;
                if (data_all) { memcpy(&v, &d[i * VW], sizeof(Vec)); reinterpret_cast<Vec *>(cw)[i] = v; }
                memcpy(&v, &p[i * VW], sizeof(Vec));
                reinterpret_cast<Vec *>(cw)[NVEC + i] = v;
            }
        } else {
#pragma unroll
            for (int w = 0; w < KW; w++) {
                if (data_all) {
                    cw[4 * w] = (uint8_t)d[w]; cw[4 * w + 1] = (uint8_t)(d[w] >> 8);
                    cw[4 * w + 2] = (uint8_t)(d[w] >> 16); cw[4 * w + 3] = (uint8_t)(d[w] >> 24);
                }
                uint8_t *o = cw + KW * 4 + 4 * w;
                o[0] = (uint8_t)p[w]; o[1] = (uint8_t)(p[w] >> 8); o[2] = (uint8_t)(p[w] >> 16); o[3] = (uint8_t)(p[w] >> 24);
            }
        }
    }
}

template <int KW, int GB>
cudaError_t launch_encode_tc_lut(DeviceCtx &ctx, const DeviceCode &dc, const uint8_t *data, uint8_t *codewords,
                                 size_t batch, cudaStream_t stream) {
    constexpr int threads = 512;
    const size_t smem = (size_t)(KW * 32 / GB) * (1 << GB) * KW * 4;
    auto kern = encode_tc_lut_kernel<KW, GB>;
    static bool configured[kMaxDevices] = {};
    static int per_sm_cached[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
        if (e != cudaSuccess) return e;
        per_sm_cached[ctx.device] = per_sm < 1 ? 1 : per_sm;
        configured[ctx.device] = true;
    }
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm_cached[ctx.device];
    const unsigned long long need = (batch + threads - 1) / threads;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, threads, smem, stream>>>(dc.enc_tc_lut, data, codewords, (unsigned long long)batch, KW * 4u);
    count_launch();
    return cudaGetLastError();
}


// TC128, 16-byte-aligned buffers: two codewords per thread, so every global access is 16 bytes wide (one load of two
// 8-byte data blocks, one store per 16-byte codeword) and twice as many bytes are in flight per thread.  Nibble rows as
// in encode_tc_lut_kernel<2, 4> (the 16 rows of a position are one sweep of the banks: no conflicts).
__global__ void __launch_bounds__(512, 2)
encode_tc128_pair_kernel(const uint32_t *__restrict__ lut_g, const uint8_t *__restrict__ data_all,
                         uint8_t *__restrict__ cw_all, unsigned long long batch) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int LUTW = 16 * 16 * 2;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(lut_g);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < LUTW / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t lut_sa = (uint32_t)__cvta_generic_to_shared(smem);
    auto parity_of = [&](uint32_t d0, uint32_t d1, uint32_t &p0, uint32_t &p1) {
        p0 = 0; p1 = 0;
#pragma unroll
        for (int w = 0; w < 2; w++) {
            const uint32_t dw = w ? d1 : d0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t byte = __byte_perm(dw, 0, 0x4440 + b);
                // high nibble = group 2j, low nibble = group 2j + 1 of byte j = 4w + b; 16 rows of 8 bytes per group
                const uint32_t a_hi = lut_sa + (uint32_t)(((4 * w + b) * 2) * 128) + (byte >> 4) * 8u;
                const uint32_t a_lo = lut_sa + (uint32_t)(((4 * w + b) * 2 + 1) * 128) + (byte & 15u) * 8u;
                uint2 x, y;
                asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x.x), "=r"(x.y) : "r"(a_hi));
                asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(y.x), "=r"(y.y) : "r"(a_lo));
                p0 ^= x.x ^ y.x; p1 ^= x.y ^ y.y;
            }
        }
    };
    const unsigned long long pairs = batch / 2;
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    auto fetch = [&](unsigned long long t) {                // (A.d0, A.d1, B.d0, B.d1) of codeword pair t
        if (data_all) return reinterpret_cast<const uint4 *>(data_all)[t];
        const uint4 a = reinterpret_cast<const uint4 *>(cw_all)[2 * t], b = reinterpret_cast<const uint4 *>(cw_all)[2 * t + 1];
        return make_uint4(a.x, a.y, b.x, b.y);
    };
    unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (t < pairs) nxt = fetch(t);
    for (; t < pairs; t += step) {
        uint4 *cw = reinterpret_cast<uint4 *>(cw_all) + 2 * t;
        const uint4 d = nxt;
        if (t + step < pairs) nxt = fetch(t + step);       // the next pair's data is in flight while this one is encoded
        uint32_t pa0, pa1, pb0, pb1;
        parity_of(d.x, d.y, pa0, pa1);
        parity_of(d.z, d.w, pb0, pb1);
        cw[0] = make_uint4(d.x, d.y, pa0, pa1);
        cw[1] = make_uint4(d.z, d.w, pb0, pb1);
    }
    if ((batch & 1) && blockIdx.x == 0 && threadIdx.x == 0) {            // odd batch: the last codeword on its own
        const unsigned long long f = batch - 1;
        const uint2 dd = data_all ? reinterpret_cast<const uint2 *>(data_all)[f] : reinterpret_cast<const uint2 *>(cw_all)[2 * f];
        uint32_t p0, p1;
        parity_of(dd.x, dd.y, p0, p1);
        reinterpret_cast<uint4 *>(cw_all)[f] = make_uint4(dd.x, dd.y, p0, p1);
    }
}

cudaError_t launch_encode_tc128_pair(DeviceCtx &ctx, const DeviceCode &dc, const uint8_t *data, uint8_t *codewords,
                                     size_t batch, cudaStream_t stream) {
    constexpr int threads = 512;
    const size_t smem = 16 * 16 * 2 * 4;
    static bool configured[kMaxDevices] = {};
    static int per_sm_cached[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        int per_sm = 1;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, encode_tc128_pair_kernel, threads, smem);
        if (e != cudaSuccess) return e;
        per_sm_cached[ctx.device] = per_sm < 1 ? 1 : per_sm;
        configured[ctx.device] = true;
    }
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm_cached[ctx.device];
    const unsigned long long need = (batch / 2 + threads - 1) / threads;
    if (grid > need) grid = need;
    if (grid == 0) grid = 1;
    encode_tc128_pair_kernel<<<(unsigned)grid, threads, smem, stream>>>(dc.enc_tc_lut, data, codewords, (unsigned long long)batch);
    count_launch();
    return cudaGetLastError();
}

// TC256 / TC512 (b = 32 / 64: a circulant block is BW = 1 / 2 whole words): byte rows of block position 0 only; a data
// byte at byte position y of its block row contributes the row of position 0 with every parity block rotated by y
// bytes (one PRMT per word) -- k/8 lookups over a 16 / 32 KB table (code_tables.h: tc_rot_encoder_lut) instead of a
// 64 KB byte table / k/4 lookups with nibble rows.
template <int BW> __host__ __device__ constexpr uint32_t rot_sel(int y, int half) {
    // PRMT selector: result byte q of word `half` of a block <- block byte (q + 4 half - y) mod (4 BW)
    uint32_t s = 0;
    for (int q = 0; q < 4; q++) s |= (uint32_t)((q + 4 * half - y) & (4 * BW - 1)) << (4 * q);
    return s;
}

// R = copies of the table in shared memory: lane l reads copy l mod R, whose rows sit R rows apart, so the lanes of a
// quarter warp (one 128-byte wavefront of a 16-byte-per-lane load) fall into different banks whatever rows their data
// select (R = 8 with 16-byte rows: never a conflict).  Costs R times the fill, so only large batches use R > 1.
template <int KW, int R>
__global__ void __launch_bounds__(R > 1 ? 1024 : 512, R > 1 ? 1 : 2)
encode_tc_rot_kernel(const uint32_t *__restrict__ lut_g, const uint8_t *__restrict__ data_all, uint8_t *__restrict__ cw_all,
                     unsigned long long batch, const uint32_t row_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int LUTW = 4 * 256 * KW, BW = KW / 4, BB = 4 * BW, NVEC = KW / 4;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(lut_g);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = threadIdx.x; i < LUTW / 4; i += blockDim.x) {
            const uint4 v = src[i];
            const int row = i / NVEC, part = i % NVEC;
#pragma unroll
            for (int g = 0; g < R; g++) dst[(row * R + g) * NVEC + part] = v;
        }
    }
    __syncthreads();
    const uint32_t lut_sa = (uint32_t)__cvta_generic_to_shared(smem) + (threadIdx.x % R) * (KW * 4);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(cw_all) | reinterpret_cast<uintptr_t>(data_all)) & 15u) == 0;
    const unsigned long long in_stride = data_all ? KW * 4ull : KW * 8ull;
    const uint8_t *in_base = data_all ? data_all : cw_all;

    // TC256: the data of the thread's next codeword is fetched before the current one is encoded (the kernel is otherwise
    // bound by the latency of these loads; +4 %).  TC512 has no registers to spare for it (it spills and loses 6 %).
    constexpr bool kPrefetch = KW == 4;
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint4 nxt[NVEC];
    if (kPrefetch && vec_ok && f < batch) {
#pragma unroll
        for (int i = 0; i < NVEC; i++) nxt[i] = reinterpret_cast<const uint4 *>(in_base + f * in_stride)[i];
    }
    for (; f < batch; f += step) {
        const uint8_t *in = in_base + f * in_stride;
        uint8_t *cw = cw_all + f * (KW * 8ull);
        uint32_t d[KW], p[KW];
        if (vec_ok) {
            if (!kPrefetch) {
#pragma unroll
                for (int i = 0; i < NVEC; i++) nxt[i] = reinterpret_cast<const uint4 *>(in)[i];
            }
#pragma unroll
            for (int i = 0; i < NVEC; i++) {
                d[4 * i] = nxt[i].x; d[4 * i + 1] = nxt[i].y; d[4 * i + 2] = nxt[i].z; d[4 * i + 3] = nxt[i].w;
            }
            if (kPrefetch && f + step < batch) {
#pragma unroll
                for (int i = 0; i < NVEC; i++) nxt[i] = reinterpret_cast<const uint4 *>(in_base + (f + step) * in_stride)[i];
            }
        } else {
#pragma unroll
            for (int w = 0; w < KW; w++)
                d[w] = (uint32_t)in[4 * w] | ((uint32_t)in[4 * w + 1] << 8) | ((uint32_t)in[4 * w + 2] << 16) | ((uint32_t)in[4 * w + 3] << 24);
        }
#pragma unroll
        for (int w = 0; w < KW; w++) p[w] = 0;
#pragma unroll
        for (int w = 0; w < KW; w++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int j = 4 * w + b, crow = j / BB, y = j % BB;         // compile-time after unrolling
                const uint32_t byte = __byte_perm(d[w], 0, 0x4440 + b);
                const uint32_t a = lut_sa + (uint32_t)(crow * 256 * KW * 4 * R) + byte * row_bytes;     // row_bytes = 4 KW R
                uint32_t t[KW];
#pragma unroll
                for (int i = 0; i < NVEC; i++)
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                        : "=r"(t[4 * i]), "=r"(t[4 * i + 1]), "=r"(t[4 * i + 2]), "=r"(t[4 * i + 3]) : "r"(a + 16 * i));
#pragma unroll
                for (int blk = 0; blk < 4; blk++) {
                    if (y == 0) {
#pragma unroll
                        for (int h = 0; h < BW; h++) p[BW * blk + h] ^= t[BW * blk + h];
                    } else if constexpr (BW == 1) {
                        p[blk] ^= __byte_perm(t[blk], t[blk], rot_sel<1>(y, 0));
                    } else {
                        p[2 * blk] ^= __byte_perm(t[2 * blk], t[2 * blk + 1], rot_sel<2>(y, 0));
                        p[2 * blk + 1] ^= __byte_perm(t[2 * blk], t[2 * blk + 1], rot_sel<2>(y, 1));
                    }
                }
            }
        }
        if (vec_ok) {
#pragma unroll
            for (int i = 0; i < NVEC; i++) {
                if (data_all) reinterpret_cast<uint4 *>(cw)[i] = make_uint4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
                reinterpret_cast<uint4 *>(cw)[NVEC + i] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
            }
        } else {
#pragma unroll
            for (int w = 0; w < KW; w++) {
                if (data_all) {
                    cw[4 * w] = (uint8_t)d[w]; cw[4 * w + 1] = (uint8_t)(d[w] >> 8);
                    cw[4 * w + 2] = (uint8_t)(d[w] >> 16); cw[4 * w + 3] = (uint8_t)(d[w] >> 24);
                }
                uint8_t *o = cw + KW * 4 + 4 * w;
                o[0] = (uint8_t)p[w]; o[1] = (uint8_t)(p[w] >> 8); o[2] = (uint8_t)(p[w] >> 16); o[3] = (uint8_t)(p[w] >> 24);
            }
        }
    }
}

template <int KW, int R>
cudaError_t launch_encode_tc_rot_r(DeviceCtx &ctx, const DeviceCode &dc, const uint8_t *data, uint8_t *codewords, size_t batch,
                                   cudaStream_t stream) {
    constexpr int threads = R > 1 ? 1024 : 512;
    const size_t smem = (size_t)4 * 256 * KW * 4 * R;
    auto kern = encode_tc_rot_kernel<KW, R>;
    static bool configured[kMaxDevices] = {};
    static int per_sm_cached[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
        if (e != cudaSuccess) return e;
        per_sm_cached[ctx.device] = per_sm < 1 ? 1 : per_sm;
        configured[ctx.device] = true;
    }
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm_cached[ctx.device];
    const unsigned long long need = (batch + threads - 1) / threads;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, threads, smem, stream>>>(dc.enc_tc_lut, data, codewords, (unsigned long long)batch, KW * 4u * R);
    count_launch();
    return cudaGetLastError();
}

template <int KW>
cudaError_t launch_encode_tc_rot(DeviceCtx &ctx, const DeviceCode &dc, const uint8_t *data, uint8_t *codewords, size_t batch,
                                 cudaStream_t stream) {
    // Table copies (128 KB per CTA): TC256 (8 copies, no conflicts at all) costs nothing measurable even for one
    // codeword; TC512 (4 copies) pays from about 128 Ki codewords (tools/enc_crossover.py).
    // LABRADOR_LDPC_ENC_TC_COPIES=0 / 1: never / always.
    static const int forced = [] { const char *e = getenv("LABRADOR_LDPC_ENC_TC_COPIES"); return e ? atoi(e) : -1; }();
    const bool copies = forced >= 0 ? forced != 0 : (KW == 4 || batch >= (1u << 17));
    if (copies) return launch_encode_tc_rot_r<KW, KW == 4 ? 8 : 4>(ctx, dc, data, codewords, batch, stream);
    return launch_encode_tc_rot_r<KW, 1>(ctx, dc, data, codewords, batch, stream);
}

template <int WPT>
cudaError_t launch_encode_wpt(DeviceCtx &ctx, const DeviceCode &dc, const uint8_t *data, uint8_t *codewords,
                              size_t batch, cudaStream_t stream) {
    const int W = (dc.n - dc.k) / 32, TPC = W / WPT;
    int threads = 256;
    if (TPC > threads) threads = TPC;
    const int fpc = threads / TPC;
    const size_t smem = ((size_t)(dc.k / dc.b) * W + (size_t)fpc * (dc.k / 32)) * sizeof(uint32_t);
    cudaError_t err = cudaFuncSetAttribute(encode_kernel<WPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (err != cudaSuccess) return err;
    unsigned long long groups = (batch + fpc - 1) / fpc;
    unsigned long long grid = groups;
    const unsigned long long cap = (unsigned long long)ctx.sm_count * 8;
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    encode_kernel<WPT><<<(unsigned)grid, threads, smem, stream>>>(dc, data, codewords, (unsigned long long)batch, fpc);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

bool launch_encode_tm(DeviceCtx &ctx, int code, const uint8_t *data, uint8_t *codewords, size_t batch,
                      cudaStream_t stream, cudaError_t *err);   // encode_tm.cu

cudaError_t launch_encode(DeviceCtx &ctx, int code, const uint8_t *data, uint8_t *codewords, size_t batch,
                          cudaStream_t stream) {
    const DeviceCode &dc = ctx.codes[code];
    if (data == codewords) data = nullptr;
    if (batch == 0) return cudaSuccess;
    // TM codes: through the sparse parity-check matrix (encode_tm.cu), 4-16x fewer bit operations than the generator;
    // LABRADOR_LDPC_ENC_GENERATOR=1 keeps the generator kernel below for A/B runs and tests
    static const bool force_gen = [] { const char *e = getenv("LABRADOR_LDPC_ENC_GENERATOR"); return e && atoi(e) != 0; }();
    if (!force_gen || !dc.gen) {
        cudaError_t err = cudaSuccess;
        if (launch_encode_tm(ctx, code, data, codewords, batch, stream, &err)) return err;
    }
    if (!dc.gen) return cudaErrorNotSupported;      // k = 16384 codes: no generator exists, the sparse-H encoder is the only one
    // TC codes: one codeword per thread over a nibble / byte lookup table, at every batch size (filling the table is
    // hidden by the launch: tools/enc_crossover.py).  The generator kernel below remains as the A/B reference.
    if (!force_gen && code < 3 && dc.enc_tc_lut) {
        // group sizes: code_tables.h: tc_encoder_group_bits
        switch (code) {
            case 0:
                if (((reinterpret_cast<uintptr_t>(codewords) | reinterpret_cast<uintptr_t>(data)) & 15u) == 0)
                    return launch_encode_tc128_pair(ctx, dc, data, codewords, batch, stream);
                return launch_encode_tc_lut<2, 4>(ctx, dc, data, codewords, batch, stream);
            case 1: return launch_encode_tc_rot<4>(ctx, dc, data, codewords, batch, stream);
            default: return launch_encode_tc_rot<8>(ctx, dc, data, codewords, batch, stream);
        }
    }
    // words per thread: the number of circulant blocks per row (n-k)/b must be divisible by it
    static const int forced = [] { const char *e = getenv("LABRADOR_LDPC_ENC_WPT"); return e ? atoi(e) : 0; }();
    const int blocks = (dc.n - dc.k) / dc.b;
    int wpt = forced ? forced : (dc.b >= 32 ? (blocks % 8 == 0 ? 8 : 4) : 1);   // measured: profiles/r01_sweep.md
    if (dc.b < 32 || blocks % wpt != 0) wpt = 1;
    // a handful of frames (the single-codeword API): latency matters, not instruction count -- one parity word per
    // thread puts 8x as many threads on each frame
    if (!forced && batch * (size_t)((dc.n - dc.k) / 32 / wpt) < (size_t)ctx.sm_count * 32) wpt = 1;
    switch (wpt) {
        case 8: return launch_encode_wpt<8>(ctx, dc, data, codewords, batch, stream);
        case 4: return launch_encode_wpt<4>(ctx, dc, data, codewords, batch, stream);
        case 2: return launch_encode_wpt<2>(ctx, dc, data, codewords, batch, stream);
        default: return launch_encode_wpt<1>(ctx, dc, data, codewords, batch, stream);
    }
}

}  // namespace ldpc
