// Batched systematic encoder: bit-packed GF(2) XOR over the compact generator.
//
// Replaces EncodeInto::encode_parity for u8/u32/u64 and LDPCCode::encode /
// copy_encode (reference src/encoder.rs:41-82, 107-160, 189-252, 292-315).
//
// The reference XORs generator row `crow` into the parity for every set data
// bit crow*b + o and rotates each b-bit parity block left once per offset, so
// the row contributed by data bit (crow, o) is the compact row gc[crow] with
// every b-bit block rotated RIGHT by o.  Here one thread owns one 32-bit
// parity word (MSB = lowest bit index, the reference's byte order) and, for
// every set data bit, XORs in the matching 32-bit window of the rotated row,
// taken with one funnel shift from two adjacent words of the row's block.
// The compact generator (<= 4 KB) and the frame's data words sit in shared
// memory.  No tensor cores: this is GF(2), not a real-valued contraction.
#include <cuda_runtime.h>

#include <cstdlib>

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
    if (!force_gen) {
        cudaError_t err = cudaSuccess;
        if (launch_encode_tm(ctx, code, data, codewords, batch, stream, &err)) return err;
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
