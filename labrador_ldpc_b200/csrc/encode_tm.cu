// Batched systematic encoder for the TM codes through the SPARSE parity-check matrix.
//
// Replaces EncodeInto::encode_parity / LDPCCode::encode / copy_encode (reference src/encoder.rs:41-82,
// 107-160, 189-252, 292-315) for TM1280 ... TM8192.  The reference multiplies the data by the dense
// generator: k * (n-k) bit operations per codeword (16.8 M for TM8192).  The codeword is the unique
// solution of H c = 0 with the given data bits, so it can equally be computed from H, which is sparse
// except for ONE M x M inverse (code_tables.h: tm_encoder_table):
//     t1, t2 = data terms of check rows 1 and 2              (sparse: windows of the data words)
//     s      = t2 + S t1,  S = the row-2 blocks of column CB  (sparse)
//     p_CC   = A^-1 s,     A = I + S G, G = the row-1 blocks of column CC   (dense, M^2 bit operations)
//     p_CB   = t1 + G p_CC,   p_CA = (I + P0) p_CC            (sparse)
// (CA, CB = the two transmitted parity block columns, CC = the punctured one.)  M^2 is 4x (rate 1/2),
// 8x (rate 2/3) and 16x (rate 4/5) less than k * (n-k).  The result is bit-identical to the generator
// encoder (tests/test_gpu_parity.py compares both with the oracle and with each other).
//
// Layout: as in decode_bf_tm.cu -- one codeword per group of min(32, M/32) lanes, bit-packed (bit i of word
// w = element 32 w + i), each lane owning one word (M = 2048: two) of every block column; identity blocks
// connect a lane's own words, pi_k blocks read a 32-bit window lo[i] : hi[i] from shared memory.  A^-1 is
// a 4 x 4 array of Q x Q circulants.  Two forms of the dense product:
//   * LUT (the default): a nibble lookup table in shared memory ("four Russians", code_tables.h:
//     tm_encoder_lut, 8 ... 128 KB; M <= 512: one interleaved copy per codeword of the warp, 64 KB): every nibble of s selects one pre-combined, pre-shifted row of M bits that is
//     XORed into the codeword's words of p_CC -- one LDS + one LOP3 per word, no branches, no bit scans;
//   * compact (A/B reference, LABRADOR_LDPC_ENC_TM_FORM=1): the 16 first columns (<= 1 KB); every set bit y of s
//     XORs the window of the column rotated by y (one funnel shift); bound by the XU pipe (BREV + FLO per bit).
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>
#include <vector>

#include "runtime.h"
#include "tm_common.cuh"

namespace ldpc {
using namespace tm;

namespace {

constexpr int kEncWarps = 8;       // compact form; the LUT form sizes its CTA from the table (launch_enc_tm)

template <class P> __host__ __device__ constexpr int n_blocks_at(int r, int c) {
    int n = 0;
    for (int b = 0; b < P::NB; b++) n += (P::blk(b).row == r && P::blk(b).col == c);
    return n;
}
template <class P> __host__ __device__ constexpr int n_pblocks_at(int r, int c) {
    int n = 0;
    for (int b = 0; b < P::NB; b++) n += (P::blk(b).row == r && P::blk(b).col == c && P::blk(b).isp);
    return n;
}

// shared-memory slot of a block column: data columns first, then CB (holds t1), then CC (holds p_CC)
template <class P> __host__ __device__ constexpr int slot_of(int col) {
    return col < P::NCOL - 3 ? col : (col == P::NCOL - 2 ? P::NCOL - 3 : P::NCOL - 2);
}

template <class P, int M> __host__ __device__ constexpr int enc_cw_stride() {
    constexpr int MW = M / 32, LPC = MW < 32 ? MW : 32;
    constexpr int words = 2 * (P::NCOL - 1) * MW;      // lo[] + hi[] of KC + 2 slots
    return LPC == 32 ? words : ((words + 31) / 32) * 32 + LPC;
}
// A^-1 first columns: [qi][qj][QW] with 5 QW words per qi so the four quarters fall into disjoint banks
template <int M> __host__ __device__ constexpr int enc_tab_words() { return ((4 * 5 * (M / 128) + 31) / 32) * 32; }

__device__ __forceinline__ uint32_t load_be32(const uint8_t *p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}

template <int IMM> __device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
    return v;
}

// Codes with several codewords per warp (M <= 512) keep one copy of the table per codeword of the warp, interleaved word
// by word: word w of a row sits at (row * MW + w) * CWW + g for copy g, so the 32 lanes of a warp -- MW words of CWW
// different rows -- always fall into 32 different banks, whatever values the codewords' nibbles have (64 KB for every
// M <= 512).  Words between the rows of consecutive nibble values:
template <int M> __host__ __device__ constexpr int enc_lut_copies() { return M / 32 < 32 ? 32 / (M / 32) : 1; }
template <int M> __host__ __device__ constexpr int enc_lut_vstride() { return 32 * (M / 32) * enc_lut_copies<M>(); }
template <int M> __host__ __device__ constexpr int enc_lut_words() { return 16 * enc_lut_vstride<M>(); }
template <int M> __host__ __device__ constexpr int enc_lut_warps() { return M >= 2048 ? 24 : (M >= 512 ? 16 : 8); }

template <int RATE, int M, bool LUT>
__global__ void __launch_bounds__(32 * (LUT ? enc_lut_warps<M>() : kEncWarps))
encode_tm_kernel(const TmParams prm, const uint32_t *__restrict__ ainv, const uint8_t *__restrict__ data_all,
                 uint8_t *__restrict__ cw_all, unsigned long long batch, const uint32_t vs_lo, const uint32_t vs_hi) {
    typedef Proto<RATE> P;
    constexpr int NB = P::NB, NCOL = P::NCOL;
    constexpr int KC = NCOL - 3, CA = NCOL - 3, CB = NCOL - 2, CC = NCOL - 1;
    constexpr int NP = count_p<P>(NB);
    constexpr int Q = M / 4, QW = Q / 32, MW = M / 32;
    constexpr int LPC = MW < 32 ? MW : 32;                // lanes per codeword
    constexpr int CWW = 32 / LPC;                         // codewords per warp
    constexpr int WPL = MW / LPC;                         // words per lane per block column
    constexpr int SLOTW = (KC + 2) * MW;
    constexpr int STRIDE = enc_cw_stride<P, M>();
    constexpr int TABW = LUT ? enc_lut_words<M>() : enc_tab_words<M>();
    constexpr int kEncWarps = LUT ? enc_lut_warps<M>() : ldpc::kEncWarps;
    constexpr unsigned kFull = 0xFFFFFFFFu;
    static_assert(Q % 32 == 0 && (QW & (QW - 1)) == 0 && MW % LPC == 0, "quarters are whole words");
    // the structure the derivation above rests on (code_tables.cpp checks the same on the run-time tables)
    static_assert(n_blocks_at<P>(0, CA) == 1 && n_pblocks_at<P>(0, CA) == 0 && n_blocks_at<P>(1, CA) == 0 &&
                  n_blocks_at<P>(2, CA) == 0, "column CA: identity in row 0 only");
    static_assert(n_blocks_at<P>(0, CB) == 0 && n_blocks_at<P>(1, CB) == 1 && n_pblocks_at<P>(1, CB) == 0 &&
                  n_blocks_at<P>(2, CB) == n_pblocks_at<P>(2, CB), "column CB: identity in row 1, pi_k blocks in row 2");
    static_assert(n_blocks_at<P>(0, CC) == 2 && n_pblocks_at<P>(0, CC) == 1 &&
                  n_blocks_at<P>(1, CC) == n_pblocks_at<P>(1, CC) && n_blocks_at<P>(2, CC) == 1 &&
                  n_pblocks_at<P>(2, CC) == 0, "column CC: I + pi in row 0, pi_k blocks in row 1, identity in row 2");

    extern __shared__ __align__(16) uint32_t smem_enc[];
    uint32_t *tab = smem_enc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LPC, wl = lane % LPC;
    uint32_t *lo = smem_enc + TABW + (warp * CWW + grp) * STRIDE;
    uint32_t *hi = lo + SLOTW;

    if constexpr (LUT) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ainv);
        uint4 *dst = reinterpret_cast<uint4 *>(tab);
        if constexpr (CWW == 1) {
            for (int i = threadIdx.x; i < TABW / 4; i += blockDim.x) dst[i] = src[i];
        } else {
            for (int i = threadIdx.x; i < 512 * MW / 4; i += blockDim.x) {
                const uint4 v = src[i];
                const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int g = 0; g < CWW; g++) tab[(4 * i + k) * CWW + g] = w4[k];
            }
        }
    } else {
        for (int i = threadIdx.x; i < 16 * QW; i += blockDim.x) {
            const int qi = i / (4 * QW);
            tab[qi * 5 * QW + (i - qi * 4 * QW)] = ainv[i];
        }
    }
    __syncthreads();

    // windows of this lane's check words into the variable bits of each pi_k block: (word index << 5) | bit shift
    uint32_t win[NP > 0 ? NP : 1][WPL];
#pragma unroll
    for (int wi = 0; wi < WPL; wi++) {
        const int e0 = (wl + wi * LPC) * 32, qa = e0 / Q, off = e0 % Q;
        static_for<0, NB>([&](auto bi) {
            constexpr int b = decltype(bi)::value;
            if constexpr (P::blk(b).isp) {
                const int qv = ((int)prm.theta[b] + qa) & 3;
                const int s0 = ((int)prm.phi[b][qa] + off) & (Q - 1);
                win[count_p<P>(b)][wi] = (uint32_t)(((slot_of<P>(P::blk(b).col) * MW + qv * QW + (s0 >> 5)) << 5) | (s0 & 31));
            }
        });
    }
    auto store_word = [&](int base, int w, uint32_t v) {
        lo[base + w] = v;
        hi[base + ((w & ~(QW - 1)) | ((w - 1) & (QW - 1)))] = v;
    };
    auto window = [&](uint32_t t) {
        const uint32_t i = t >> 5;
        return __funnelshift_r(lo[i], hi[i], t);
    };

    const unsigned long long n_groups = (batch + CWW - 1) / CWW;
    const bool aligned = ((reinterpret_cast<uintptr_t>(cw_all) | reinterpret_cast<uintptr_t>(data_all)) & 3u) == 0;
    const unsigned long long in_stride = data_all ? (unsigned long long)(KC * MW * 4) : (unsigned long long)((KC + 2) * MW * 4);
    const uint8_t *in_base = data_all ? data_all : cw_all;

    // raw (little-endian) data words of codeword group g; the next group's are fetched while this one is encoded
    auto fetch = [&](unsigned long long g, uint32_t (&raw)[KC][WPL]) {
        const unsigned long long first = g * CWW, frame = first + grp;
        const uint8_t *in = in_base + (frame < batch ? frame : first) * in_stride;
#pragma unroll
        for (int wi = 0; wi < WPL; wi++)
#pragma unroll
            for (int c = 0; c < KC; c++) {
                const int iw = c * MW + wl + wi * LPC;
                if (aligned) raw[c][wi] = reinterpret_cast<const uint32_t *>(in)[iw];
                else raw[c][wi] = __byte_perm(load_be32(in + 4 * iw), 0, 0x0123);
            }
    };
    const unsigned long long g_step = (unsigned long long)gridDim.x * kEncWarps;
    unsigned long long g = (unsigned long long)blockIdx.x * kEncWarps + warp;
    uint32_t nxt[KC][WPL];
    if (g < n_groups) fetch(g, nxt);

    for (; g < n_groups; g += g_step) {
        const unsigned long long first = g * CWW;
        const unsigned long long frame = first + grp;
        const bool live = frame < batch;
        uint8_t *cw = cw_all + (live ? frame : first) * (unsigned long long)((KC + 2) * MW * 4);

        // data words (bit-reversed: bit i of word w = element 32 w + i); copy_encode also writes them out
        uint32_t dw[KC][WPL];
#pragma unroll
        for (int wi = 0; wi < WPL; wi++) {
            const int w = wl + wi * LPC;
#pragma unroll
            for (int c = 0; c < KC; c++) {
                const int iw = c * MW + w;
                const uint32_t raw = nxt[c][wi];
                if (data_all && live) {
                    if (aligned) {
                        reinterpret_cast<uint32_t *>(cw)[iw] = raw;
                    } else {
                        cw[4 * iw + 0] = (uint8_t)raw; cw[4 * iw + 1] = (uint8_t)(raw >> 8);
                        cw[4 * iw + 2] = (uint8_t)(raw >> 16);  cw[4 * iw + 3] = (uint8_t)(raw >> 24);
                    }
                }
                dw[c][wi] = __brev(__byte_perm(raw, 0, 0x0123));
                store_word(c * MW, w, dw[c][wi]);
            }
        }
        if (g + g_step < n_groups) fetch(g + g_step, nxt);
        __syncwarp();

        // t1 / t2: data terms of check rows 1 and 2
        uint32_t t1[WPL], sv[WPL];
#pragma unroll
        for (int wi = 0; wi < WPL; wi++) {
            uint32_t x1 = 0, x2 = 0;
            static_for<0, NB>([&](auto bi) {
                constexpr int b = decltype(bi)::value;
                if constexpr (P::blk(b).col < KC) {
                    uint32_t v;
                    if constexpr (P::blk(b).isp) v = window(win[count_p<P>(b)][wi]);
                    else v = dw[P::blk(b).col][wi];
                    if constexpr (P::blk(b).row == 1) x1 ^= v; else x2 ^= v;
                }
            });
            t1[wi] = x1; sv[wi] = x2;
            store_word(KC * MW, wl + wi * LPC, x1);          // t1 stands in for column CB
        }
        __syncwarp();
        // s = t2 + S t1
#pragma unroll
        for (int wi = 0; wi < WPL; wi++) {
            static_for<0, NB>([&](auto bi) {
                constexpr int b = decltype(bi)::value;
                if constexpr (P::blk(b).col == CB && P::blk(b).row == 2) sv[wi] ^= window(win[count_p<P>(b)][wi]);
            });
        }

        // p_CC = A^-1 s:  p[qi Q + x] = XOR over qj, y of col_{qi,qj}[(x - y) mod Q] s[qj Q + y]
        uint32_t pc[WPL];
#pragma unroll
        for (int wi = 0; wi < WPL; wi++) pc[wi] = 0;
        if constexpr (LUT) {
            // Row of the table for (nibble value v, source quarter qj, nibble position nib): byte offset
            // v * VS + (qj * 8 + nib) * 4 MW CWW (VS = enc_lut_vstride words).  v * VS comes from one LOP3 (the nibble in place inside its byte) and
            // one IMAD whose multiplier is a kernel parameter (keeps it on the FMA pipe); the rest is an immediate.
            const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(tab);
#pragma unroll
            for (int wj = 0; wj < WPL; wj++) {
#pragma unroll 2
                for (int l = 0; l < LPC; l++) {
                    const uint32_t D = __shfl_sync(kFull, sv[wj], grp * LPC + l);
                    const int j = l + wj * LPC, qj = j / QW, wq = j % QW;
                    // this lane's first word inside a row; its other words follow LPC words apart
                    const uint32_t a0 = tab_sa + (uint32_t)(((qj * 8 * MW + ((wl & ~(QW - 1)) | ((wl - wq) & (QW - 1)))) * CWW + grp) * 4);
                    static_for<0, 4>([&](auto bi) {
                        constexpr int b = decltype(bi)::value;
                        const uint32_t Db = __byte_perm(D, 0, 0x4440 + b);
                        const uint32_t r0 = (Db & 0x0Fu) * vs_lo + a0;          // vs_lo = VS, vs_hi = VS / 16
                        const uint32_t r1 = (Db & 0xF0u) * vs_hi + a0;
                        static_for<0, WPL>([&](auto wii) {
                            constexpr int wi = decltype(wii)::value;
                            pc[wi] ^= lds_u32<((2 * b) * MW + wi * LPC) * CWW * 4>(r0) ^ lds_u32<((2 * b + 1) * MW + wi * LPC) * CWW * 4>(r1);
                        });
                    });
                }
            }
        } else {
    #pragma unroll
            for (int wj = 0; wj < WPL; wj++) {
    #pragma unroll 2
                for (int l = 0; l < LPC; l++) {
                    uint32_t D = __shfl_sync(kFull, sv[wj], grp * LPC + l);
                    const int j = l + wj * LPC, qj = j / QW, wq = j % QW;
                    uint32_t xl[WPL], xh[WPL];
    #pragma unroll
                    for (int wi = 0; wi < WPL; wi++) {
                        const int w = wl + wi * LPC, qi = w / QW, wx = w % QW;
                        const int w0 = (wx - wq) & (QW - 1);
                        const uint32_t *col = tab + qi * 5 * QW + qj * QW;
                        xh[wi] = col[w0];
                        xl[wi] = col[(w0 - 1) & (QW - 1)];
                    }
                    while (D) {
                        const int o = __ffs((int)D) - 1;
                        D &= D - 1;
    #pragma unroll
                        for (int wi = 0; wi < WPL; wi++) pc[wi] ^= __funnelshift_l(xl[wi], xh[wi], o);
                    }
                }
            }
        }
#pragma unroll
        for (int wi = 0; wi < WPL; wi++) store_word((KC + 1) * MW, wl + wi * LPC, pc[wi]);
        __syncwarp();

        // p_CB = t1 + G p_CC,  p_CA = (I + P0) p_CC
#pragma unroll
        for (int wi = 0; wi < WPL; wi++) {
            uint32_t pa = 0, pb = t1[wi];
            static_for<0, NB>([&](auto bi) {
                constexpr int b = decltype(bi)::value;
                if constexpr (P::blk(b).col == CC && P::blk(b).row == 1) pb ^= window(win[count_p<P>(b)][wi]);
                if constexpr (P::blk(b).col == CC && P::blk(b).row == 0) {
                    if constexpr (P::blk(b).isp) pa ^= window(win[count_p<P>(b)][wi]);
                    else pa ^= pc[wi];
                }
            });
            if (live) {
                const int w = wl + wi * LPC;
                const uint32_t ra = __brev(pa), rb = __brev(pb);
                if (aligned) {
                    reinterpret_cast<uint32_t *>(cw)[KC * MW + w] = __byte_perm(ra, 0, 0x0123);
                    reinterpret_cast<uint32_t *>(cw)[(KC + 1) * MW + w] = __byte_perm(rb, 0, 0x0123);
                } else {
                    uint8_t *o = cw + 4 * (KC * MW + w);
                    o[0] = (uint8_t)(ra >> 24); o[1] = (uint8_t)(ra >> 16); o[2] = (uint8_t)(ra >> 8); o[3] = (uint8_t)ra;
                    o = cw + 4 * ((KC + 1) * MW + w);
                    o[0] = (uint8_t)(rb >> 24); o[1] = (uint8_t)(rb >> 16); o[2] = (uint8_t)(rb >> 8); o[3] = (uint8_t)rb;
                }
            }
        }
        __syncwarp();      // the next codeword overwrites the slots
    }
}

template <int RATE, int M, bool LUT>
cudaError_t launch_enc_tm_form(DeviceCtx &ctx, const CodeInfo &c, const DeviceCode &dc, const uint8_t *data,
                               uint8_t *codewords, size_t batch, cudaStream_t stream) {
    typedef Proto<RATE> P;
    const TmParams prm = make_params<RATE>(c);
    constexpr int MW = M / 32, CWW = MW < 32 ? 32 / MW : 1;
    constexpr int warps = LUT ? enc_lut_warps<M>() : kEncWarps;
    constexpr int tabw = LUT ? enc_lut_words<M>() : enc_tab_words<M>();
    const size_t smem = ((size_t)tabw + (size_t)warps * CWW * enc_cw_stride<P, M>()) * sizeof(uint32_t);
    auto kern = encode_tm_kernel<RATE, M, LUT>;
    static bool configured[kMaxDevices] = {};
    static int per_sm_cached[kMaxDevices] = {};
    if (!configured[ctx.device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * warps, smem);
        if (e != cudaSuccess) return e;
        per_sm_cached[ctx.device] = per_sm < 1 ? 1 : per_sm;
        configured[ctx.device] = true;
    }
    unsigned long long grid = (unsigned long long)ctx.sm_count * per_sm_cached[ctx.device];
    const unsigned long long groups = (batch + CWW - 1) / CWW;
    const unsigned long long need = (groups + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid == 0) grid = 1;
    kern<<<(unsigned)grid, 32 * warps, smem, stream>>>(prm, LUT ? dc.enc_lut : dc.enc_ainv, data, codewords,
                                                       (unsigned long long)batch, enc_lut_vstride<M>() * 4u, enc_lut_vstride<M>() * 4u / 16u);
    count_launch();
    return cudaGetLastError();
}

template <int RATE, int M>
cudaError_t launch_enc_tm(DeviceCtx &ctx, const CodeInfo &c, const DeviceCode &dc, const uint8_t *data,
                          uint8_t *codewords, size_t batch, cudaStream_t stream) {
    // The LUT form fills 8 ... 128 KB of shared memory per CTA before the first codeword, which costs less than the
    // bit-serial dense step of the compact form even for ONE codeword (tools/enc_crossover.py: TM8192 21 vs 38 us per
    // call at 64 codewords); the compact form stays as the A/B reference.
    // LABRADOR_LDPC_ENC_TM_FORM = 1 (compact) / 2 (LUT) forces one.
    static const int forced = [] { const char *e = getenv("LABRADOR_LDPC_ENC_TM_FORM"); return e ? atoi(e) : 0; }();
    const bool lut = forced != 1;
    if (lut && dc.enc_lut) return launch_enc_tm_form<RATE, M, true>(ctx, c, dc, data, codewords, batch, stream);
    return launch_enc_tm_form<RATE, M, false>(ctx, c, dc, data, codewords, batch, stream);
}

}  // namespace

// Returns true (and launches) for the TM codes whose parity-check-based encoder table exists.
bool launch_encode_tm(DeviceCtx &ctx, int code, const uint8_t *data, uint8_t *codewords, size_t batch,
                      cudaStream_t stream, cudaError_t *err) {
    if (code < 3 || code >= kNumCodes) return false;
    const DeviceCode &dc = ctx.codes[code];
    if (!dc.enc_ainv) return false;
    const CodeInfo &c = *code_info(code);
    switch (code) {
        // the k = 16384 codes (no generator exists for them in the reference: this is their only encoder).  The nibble
        // table of M = 4096 / 8192 would not fit in shared memory, so all three use the compact form.
        case 9: if (!structure_matches<2>(c) || c.m != 2048) return false;
                *err = launch_enc_tm_form<2, 2048, false>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 10: if (!structure_matches<1>(c) || c.m != 4096) return false;
                *err = launch_enc_tm_form<1, 4096, false>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 11: if (!structure_matches<0>(c) || c.m != 8192) return false;
                *err = launch_enc_tm_form<0, 8192, false>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 3: if (!structure_matches<2>(c) || c.m != 128) return false;
                *err = launch_enc_tm<2, 128>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 4: if (!structure_matches<1>(c) || c.m != 256) return false;
                *err = launch_enc_tm<1, 256>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 5: if (!structure_matches<0>(c) || c.m != 512) return false;
                *err = launch_enc_tm<0, 512>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 6: if (!structure_matches<2>(c) || c.m != 512) return false;
                *err = launch_enc_tm<2, 512>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 7: if (!structure_matches<1>(c) || c.m != 1024) return false;
                *err = launch_enc_tm<1, 1024>(ctx, c, dc, data, codewords, batch, stream); return true;
        case 8: if (!structure_matches<0>(c) || c.m != 2048) return false;
                *err = launch_enc_tm<0, 2048>(ctx, c, dc, data, codewords, batch, stream); return true;
        default: return false;
    }
}

}  // namespace ldpc
