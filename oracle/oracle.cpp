// CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// A plain C++ restatement of the labrador-ldpc v1.2.1 hot path (encode,
// decode_ms, decode_bf, decode_erasures, hard_to_llrs, llrs_to_hard and the
// parity-check edge iterator).  It exists so the CUDA kernels can be checked
// bit-for-bit against the reference's algorithm on identical inputs.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this file's shared object.  The product library
// (labrador_ldpc_b200/csrc) never includes, links or calls anything here.
//
// Parity status: PINNED.  The reference is Rust and neither cargo nor rustc
// exist in this image, so the Rust binary itself cannot be run; this
// restatement is pinned instead by every golden the reference's own tests
// hold for the path (tests/test_oracle_goldens.py):
//   * 9 CRC-32 edge-order goldens + edge counts   src/codes/mod.rs:517-535
//   * 9 full parity-block known answers           src/encoder.rs:361-527
//   * doc-test vectors                            src/lib.rs:21-50, 130-144
//   * converter vectors                           src/decoder.rs:553-605
//   * length formulas vs literal params           src/decoder.rs:531-551
//   * behavioural decode tests                    src/decoder.rs:607-699
// What the reference's tests do NOT pin (returned iteration counts, noisy
// soft-input decodes, i16/i32/f32/f64 decodes) rests on the line-by-line
// correspondence of decode_ms() below with src/decoder.rs:347-475.
//
// Build: see oracle/Makefile (g++ -O2 -fno-fast-math -ffp-contract=off).

#include <atomic>
#include <cfloat>
#include <climits>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "ccsds_tables.h"

namespace {

// ---------------------------------------------------------------------------
// Code parameters: src/codes/mod.rs:69-241 (literal table) and :367-409.
// ---------------------------------------------------------------------------
struct Params {
    int n, k, p, m, b;
    int edges;
    const uint8_t *proto;   // [3][4][11]
    const uint16_t *phi;    // [4][26], nullptr for TC codes
    const uint64_t *gen;
};

const Params PARAMS[12] = {
    {128, 64, 0, 16, 16, 512, ccsds_proto_tc128, nullptr, ccsds_gen_tc128},
    {256, 128, 0, 32, 32, 1024, ccsds_proto_tc256, nullptr, ccsds_gen_tc256},
    {512, 256, 0, 64, 64, 2048, ccsds_proto_tc512, nullptr, ccsds_gen_tc512},
    {1280, 1024, 128, 128, 32, 4992, ccsds_proto_tm_r45, ccsds_phi_m128, ccsds_gen_tm1280},
    {1536, 1024, 256, 256, 64, 5888, ccsds_proto_tm_r23, ccsds_phi_m256, ccsds_gen_tm1536},
    {2048, 1024, 512, 512, 128, 7680, ccsds_proto_tm_r12, ccsds_phi_m512, ccsds_gen_tm2048},
    {5120, 4096, 512, 512, 128, 19968, ccsds_proto_tm_r45, ccsds_phi_m512, ccsds_gen_tm5120},
    {6144, 4096, 1024, 1024, 256, 23552, ccsds_proto_tm_r23, ccsds_phi_m1024, ccsds_gen_tm6144},
    {8192, 4096, 2048, 2048, 512, 30720, ccsds_proto_tm_r12, ccsds_phi_m2048, ccsds_gen_tm8192},
    // The k = 16384 codes.  The reference has their parity-check constants (compact_parity_checks.rs:84-96; the phi
    // tables for M = 4096 / 8192 are selected at src/codes/mod.rs:473-476) but no enum variant, parameters or
    // generators (src/lib.rs:81-83).  Parameters below follow the pattern of the six TM codes above (p = M, b = M/4,
    // edges = blocks * M); the decoders are the same code run over these parameters; there is no encoder (gen ==
    // nullptr) and NO reference golden of any kind for them -- see tests/test_gpu_k16384.py for how they are pinned.
    {20480, 16384, 2048, 2048, 512, 39 * 2048, ccsds_proto_tm_r45, ccsds_phi_m2048, nullptr},
    {24576, 16384, 4096, 4096, 1024, 23 * 4096, ccsds_proto_tm_r23, ccsds_phi_m4096, nullptr},
    {32768, 16384, 8192, 8192, 2048, 15 * 8192, ccsds_proto_tm_r12, ccsds_phi_m8192, nullptr},
};

inline bool valid(int code) { return code >= 0 && code < 12; }

// src/decoder.rs:93-116
inline size_t bf_working_len(const Params &c) { return c.n + c.p; }
inline size_t ms_working_len(const Params &c) { return 2 * (size_t)c.edges + 3 * c.n + 3 * c.p - 2 * c.k; }
inline size_t ms_working_u8_len(const Params &c) { return (c.n + c.p - c.k) / 8; }
inline size_t output_len(const Params &c) { return (c.n + c.p) / 8; }

// ---------------------------------------------------------------------------
// Edge iterator: src/codes/mod.rs:275-362 (next) and :444-494 (setup).
//
// The reference's state machine visits prototype rows 0..3, columns 0..10 and
// for each cell the summed sub-prototypes 0..2, stopping at the first zero
// entry of a cell (:332-339); every non-zero entry yields M edges with
// check = row*M + i, i = 0..M-1, and
//   identity:    var = col*M + ((i + val) mod M)                    (:305-311)
//   permutation: var = col*M + M/4*((theta[val] + i/(M/4)) mod 4)
//                            + ((phi[i/(M/4)][val] + i) mod M/4)    (:312-322)
// ---------------------------------------------------------------------------
template <class F>
inline void for_each_edge(const Params &c, F &&f) {
    const int m = c.m;
    const int q = m / 4;
    int logq = 0;
    while ((1 << logq) < q) logq++;
    int idx = 0;   // running edge index (the reference's `idx += 1`, src/decoder.rs:410,449)
    for (int row = 0; row < 4; row++) {
        for (int col = 0; col < 11; col++) {
            for (int sub = 0; sub < 3; sub++) {
                const uint8_t e = c.proto[(sub * 4 + row) * 11 + col];
                if (e == 0) break;
                const int val = e & CCSDS_VAL_MASK;
                const int kind = e & CCSDS_KIND_MASK;
                if (kind == CCSDS_KIND_IDENT) {
                    for (int i = 0; i < m; i++, idx++)
                        f(idx, row * m + i, col * m + ((i + val) & (m - 1)));
                } else if (kind == CCSDS_KIND_PERM) {
                    for (int i = 0; i < m; i++, idx++) {
                        const int j = i >> logq;
                        const int pi = (((ccsds_theta_k[val] + j) % 4) << logq) +
                                       ((c.phi[j * 26 + val] + i) & (q - 1));
                        f(idx, row * m + i, col * m + pi);
                    }
                }
            }
        }
    }
}

// src/codes/mod.rs:508-515
inline uint32_t crc32_u16(uint32_t crc, uint32_t data) {
    crc ^= data;
    for (int i = 0; i < 16; i++) {
        const uint32_t mask = (crc & 1) ? 0xFFFFFFFFu : 0u;
        crc = (crc >> 1) ^ (0xEDB88320u & mask);
    }
    return crc;
}

// ---------------------------------------------------------------------------
// Encoders: src/encoder.rs:41-82 (u8), :107-160 (u32), :189-252 (u64).
// ---------------------------------------------------------------------------
void encode_parity_u8(const Params &c, const uint8_t *data, uint8_t *parity) {
    const int k = c.k, r = c.n - c.k, b = c.b;
    const uint64_t *gc = c.gen;
    const int row_len = r / 64;
    for (int i = 0; i < r / 8; i++) parity[i] = 0;
    for (int offset = 0; offset < b; offset++) {
        for (int crow = 0; crow < k / b; crow++) {
            const int bit = crow * b + offset;
            if ((data[bit / 8] >> (7 - (bit % 8))) & 1) {
                for (int idx = 0; idx < row_len; idx++) {
                    const uint64_t circ = gc[crow * row_len + idx];
                    for (int byte = 0; byte < 8; byte++)
                        parity[idx * 8 + byte] ^= (uint8_t)(circ >> (56 - 8 * byte));
                }
            }
        }
        for (int block = 0; block < r / b; block++) {
            uint8_t *pb = parity + block * b / 8;
            const int len = b / 8;
            uint8_t carry = pb[0] >> 7;
            for (int x = len - 1; x >= 0; x--) {
                const uint8_t cc = pb[x] >> 7;
                pb[x] = (uint8_t)((pb[x] << 1) | carry);
                carry = cc;
            }
        }
    }
}

inline uint32_t to_be32(uint32_t x) { return __builtin_bswap32(x); }
inline uint64_t to_be64(uint64_t x) { return __builtin_bswap64(x); }

void encode_parity_u32(const Params &c, const uint8_t *data, uint32_t *parity) {
    const int k = c.k, r = c.n - c.k, b = c.b;
    const uint64_t *gc = c.gen;
    const int row_len = r / 64;
    const int nw = r / 32;
    for (int i = 0; i < nw; i++) parity[i] = 0;
    for (int offset = 0; offset < b; offset++) {
        for (int crow = 0; crow < k / b; crow++) {
            const int bit = crow * b + offset;
            if ((data[bit / 8] >> (7 - (bit % 8))) & 1) {
                for (int idx = 0; idx < row_len; idx++) {
                    const uint64_t circ = gc[crow * row_len + idx];
                    parity[idx * 2 + 1] ^= (uint32_t)(circ >> 0);
                    parity[idx * 2 + 0] ^= (uint32_t)(circ >> 32);
                }
            }
        }
        if (b >= 32) {
            for (int block = 0; block < r / b; block++) {
                uint32_t *pb = parity + block * b / 32;
                const int len = b / 32;
                uint32_t carry = pb[0] >> 31;
                for (int x = len - 1; x >= 0; x--) {
                    const uint32_t cc = pb[x] >> 31;
                    pb[x] = (pb[x] << 1) | carry;
                    carry = cc;
                }
            }
        } else if (b == 16) {
            for (int i = 0; i < nw; i++) {
                const uint32_t x = parity[i];
                const uint32_t b1 = x & 0xFFFF0000u, b2 = x & 0x0000FFFFu;
                parity[i] = (((b1 << 1) | (b1 >> 15)) & 0xFFFF0000u) |
                            (((b2 << 1) | (b2 >> 15)) & 0x0000FFFFu);
            }
        }
    }
    for (int i = 0; i < nw; i++) parity[i] = to_be32(parity[i]);
}

void encode_parity_u64(const Params &c, const uint8_t *data, uint64_t *parity) {
    const int k = c.k, r = c.n - c.k, b = c.b;
    const uint64_t *gc = c.gen;
    const int row_len = r / 64;
    const int nw = r / 64;
    for (int i = 0; i < nw; i++) parity[i] = 0;
    for (int offset = 0; offset < b; offset++) {
        for (int crow = 0; crow < k / b; crow++) {
            const int bit = crow * b + offset;
            if ((data[bit / 8] >> (7 - (bit % 8))) & 1) {
                for (int idx = 0; idx < row_len; idx++) parity[idx] ^= gc[crow * row_len + idx];
            }
        }
        if (b >= 64) {
            for (int block = 0; block < r / b; block++) {
                uint64_t *pb = parity + block * b / 64;
                const int len = b / 64;
                uint64_t carry = pb[0] >> 63;
                for (int x = len - 1; x >= 0; x--) {
                    const uint64_t cc = pb[x] >> 63;
                    pb[x] = (pb[x] << 1) | carry;
                    carry = cc;
                }
            }
        } else if (b == 32) {
            for (int i = 0; i < nw; i++) {
                const uint64_t x = parity[i];
                const uint64_t b1 = x & 0xFFFFFFFF00000000ull, b2 = x & 0x00000000FFFFFFFFull;
                parity[i] = (((b1 << 1) | (b1 >> 31)) & 0xFFFFFFFF00000000ull) |
                            (((b2 << 1) | (b2 >> 31)) & 0x00000000FFFFFFFFull);
            }
        } else if (b == 16) {
            for (int i = 0; i < nw; i++) {
                const uint64_t x = parity[i];
                uint64_t y = 0;
                for (int s = 0; s < 64; s += 16) {
                    const uint64_t msk = 0xFFFFull << s;
                    const uint64_t blk = x & msk;
                    y |= ((blk << 1) | (blk >> 15)) & msk;
                }
                parity[i] = y;
            }
        }
    }
    for (int i = 0; i < nw; i++) parity[i] = to_be64(parity[i]);
}

// ---------------------------------------------------------------------------
// DecodeFrom scalar semantics: src/decoder.rs:42-86.
// ---------------------------------------------------------------------------
template <class T> struct Ops;

template <class T, class W, W LO, W HI> struct IntOps {
    static T one() { return 1; }
    static T zero() { return 0; }
    static T maxval() { return (T)HI; }
    static T clampw(W x) { return (T)(x < LO ? LO : (x > HI ? HI : x)); }
    static T abs(T x) { return x < 0 ? clampw(-(W)x) : x; }              // saturating_abs
    static T sat_add(T a, T b) { return clampw((W)a + (W)b); }
    static T sat_sub(T a, T b) { return clampw((W)a - (W)b); }
    static T neg(T x) { return (T)(-x); }
    static bool hard_bit(T x) { return x < 0; }
};
template <> struct Ops<int8_t> : IntOps<int8_t, int32_t, INT8_MIN, INT8_MAX> {};
template <> struct Ops<int16_t> : IntOps<int16_t, int32_t, INT16_MIN, INT16_MAX> {};
template <> struct Ops<int32_t> : IntOps<int32_t, int64_t, INT32_MIN, INT32_MAX> {};

template <> struct Ops<float> {
    static float one() { return 1.0f; }
    static float zero() { return 0.0f; }
    static float maxval() { return FLT_MAX; }
    static float abs(float x) {
        uint32_t u; memcpy(&u, &x, 4); u &= 0x7FFFFFFFu; memcpy(&x, &u, 4); return x;
    }
    static float sat_add(float a, float b) { return a + b; }
    static float sat_sub(float a, float b) { return a - b; }
    static float neg(float x) { return -x; }
    static bool hard_bit(float x) { return x < 0.0f; }
};
template <> struct Ops<double> {
    static double one() { return 1.0; }
    static double zero() { return 0.0; }
    static double maxval() { return DBL_MAX; }
    static double abs(double x) {
        uint64_t u; memcpy(&u, &x, 8); u &= 0x7FFFFFFFFFFFFFFFull; memcpy(&x, &u, 8); return x;
    }
    static double sat_add(double a, double b) { return a + b; }
    static double sat_sub(double a, double b) { return a - b; }
    static double neg(double x) { return -x; }
    static bool hard_bit(double x) { return x < 0.0; }
};

// ---------------------------------------------------------------------------
// decode_ms: src/decoder.rs:347-475.  Same buffers, same two edge loops.
// ---------------------------------------------------------------------------
template <class T>
bool decode_ms(const Params &c, const T *llrs, uint8_t *output, T *working, uint8_t *working_u8,
               size_t maxiters, size_t *iters_run) {
    typedef Ops<T> O;
    const int n = c.n, k = c.k, p = c.p;
    const size_t E = c.edges;
    const size_t olen = output_len(c);
    const size_t nchk = n + p - k;

    uint8_t *parities = output;                                   // :363
    uint8_t *ui_sgns = working_u8;                                // :367
    for (size_t i = 0; i < ms_working_u8_len(c); i++) ui_sgns[i] = 0;     // :368
    for (size_t i = 0; i < ms_working_len(c); i++) working[i] = O::zero();  // :374
    T *u = working;                                               // :375-378
    T *v = u + E;
    T *va = v + E;
    T *ui_min1 = va + (n + p);
    T *ui_min2 = ui_min1 + nchk;

    for (size_t iter = 0; iter < maxiters; iter++) {              // :380
        for (int i = 0; i < n; i++) va[i] = llrs[i];              // :382
        for (int i = n; i < n + p; i++) va[i] = O::zero();        // :383

        for_each_edge(c, [=](int idx, int check, int var) {       // :387-388
            // (locals + conditional expressions instead of repeated u[idx] stores so the
            // compiler can emit selects, as rustc does; same operations in the same order)
            const T vi = v[idx];
            const T m1 = ui_min1[check];
            T uu = (O::abs(vi) == m1) ? ui_min2[check] : m1;              // :391-395
            uu = ((ui_sgns[check / 8] >> (check % 8)) & 1) ? O::neg(uu) : uu;  // :398-400
            uu = O::hard_bit(vi) ? O::neg(uu) : uu;                       // :403-405
            u[idx] = uu;
            va[var] = O::sat_add(va[var], uu);                            // :408
        });

        for (size_t i = 0; i < nchk; i++) ui_min1[i] = O::maxval();   // :414
        for (size_t i = 0; i < nchk; i++) ui_min2[i] = O::maxval();   // :415
        for (size_t i = 0; i < ms_working_u8_len(c); i++) ui_sgns[i] = 0;  // :416
        for (size_t i = 0; i < olen; i++) parities[i] = 0;        // :417
        for_each_edge(c, [=](int idx, int check, int var) {       // :419
            const T vav = va[var];
            const T vold = v[idx];
            const T new_v = O::sat_sub(vav, u[idx]);              // :421
            const T vn = (O::hard_bit(new_v) == O::hard_bit(vold) || vold == O::zero())
                             ? new_v : O::zero();                 // :422-426
            v[idx] = vn;
            const T a = O::abs(vn);
            const T m1 = ui_min1[check], m2 = ui_min2[check];     // :430-435
            ui_min1[check] = (a < m1) ? a : m1;
            ui_min2[check] = (a < m1) ? m1 : ((a < m2) ? a : m2);
            ui_sgns[check / 8] ^= (uint8_t)((O::hard_bit(vn) ? 1 : 0) << (check % 8));    // :439-441
            parities[check / 8] ^= (uint8_t)((O::hard_bit(vav) ? 1 : 0) << (check % 8));  // :445-447
        });

        uint8_t mx = 0;                                           // :453
        for (size_t i = 0; i < olen; i++) if (parities[i] > mx) mx = parities[i];
        if (mx == 0) {
            for (size_t i = 0; i < olen; i++) output[i] = 0;      // :456
            for (int a = 0; a < n + p; a++)
                if (O::hard_bit(va[a])) output[a / 8] |= (uint8_t)(1 << (7 - (a % 8)));
            if (iters_run) *iters_run = iter;
            return true;                                          // :462
        }
    }
    for (size_t i = 0; i < olen; i++) output[i] = 0;              // :468
    for (int a = 0; a < n + p; a++)
        if (O::hard_bit(va[a])) output[a / 8] |= (uint8_t)(1 << (7 - (a % 8)));
    if (iters_run) *iters_run = maxiters;
    return false;                                                 // :474
}

// ---------------------------------------------------------------------------
// decode_erasures: src/decoder.rs:144-223.
// ---------------------------------------------------------------------------
bool decode_erasures(const Params &c, uint8_t *codeword, uint8_t *working, size_t maxiters,
                     size_t *iters_run) {
    const int n = c.n, p = c.p;
    for (int i = 0; i < n; i++) working[i] = 0x00;                // :163
    for (int i = n; i < n + p; i++) working[i] = 0x10;            // :164
    for (size_t i = n / 8; i < output_len(c); i++) codeword[i] = 0x00;   // :167
    int bits_fixed = 0;                                           // :170
    for (size_t iter = 0; iter < maxiters; iter++) {
        for (int i = 0; i < n + p; i++) working[i] = (working[i] & 0x10) | 0x08;  // :174
        for_each_edge(c, [&](int, int check, int var) {                // :177-189
            if ((working[var] & 0x10) == 0x10) {
                switch (working[check] & 0x60) {
                    case 0x00: working[check] |= 0x20; break;
                    case 0x20: working[check] |= 0x40; break;
                    default: break;
                }
            } else if ((codeword[var / 8] >> (7 - (var % 8))) & 1) {
                working[check] ^= 0x80;
            }
        });
        for_each_edge(c, [&](int, int check, int var) {                // :192-202
            if ((working[var] & 0x10) == 0x10 && (working[check] & 0x60) == 0x20) {
                if ((working[check] & 0x80) == 0x80) working[var] += 1;
                else working[var] -= 1;
            }
        });
        for (int var = 0; var < n + p; var++) {                   // :205-213
            if ((working[var] & 0x10) == 0x10) {
                if ((working[var] & 0x0F) > 0x08) {
                    codeword[var / 8] |= (uint8_t)(1 << (7 - (var % 8)));
                    working[var] &= (uint8_t)~0x10;
                }
                bits_fixed += 1;
            }
        }
        if (bits_fixed == p) {                                    // :215-218
            if (iters_run) *iters_run = iter;
            return true;
        }
    }
    if (iters_run) *iters_run = maxiters;
    return false;                                                 // :222
}

// ---------------------------------------------------------------------------
// decode_bf: src/decoder.rs:243-301.
// ---------------------------------------------------------------------------
bool decode_bf(const Params &c, const uint8_t *input, uint8_t *output, uint8_t *working,
               size_t maxiters, size_t *iters_run) {
    const int n = c.n, p = c.p;
    memcpy(output, input, n / 8);                                 // :251
    size_t erasure_iters = 0;                                     // :256-259
    if (p > 0) decode_erasures(c, output, working, maxiters, &erasure_iters);
    for (size_t iter = 0; iter < maxiters; iter++) {              // :264
        for (int i = 0; i < n + p; i++) working[i] = 0;           // :266
        for_each_edge(c, [&](int, int check, int var) {                // :269-273
            if ((output[var / 8] >> (7 - (var % 8))) & 1) working[check] ^= 0x80;
        });
        uint8_t max_violations = 0;                               // :276-286
        for_each_edge(c, [&](int, int check, int var) {
            if ((working[check] & 0x80) == 0x80) {
                working[var] += 1;
                if ((working[var] & 0x7F) > max_violations) max_violations = working[var] & 0x7F;
            }
        });
        if (max_violations == 0) {                                // :288-289
            if (iters_run) *iters_run = iter + erasure_iters;
            return true;
        }
        for (int var = 0; var < n + p; var++)                     // :292-296
            if ((working[var] & 0x7F) == max_violations)
                output[var / 8] ^= (uint8_t)(1 << (7 - (var % 8)));
    }
    if (iters_run) *iters_run = maxiters + erasure_iters;
    return false;                                                 // :300
}

// src/decoder.rs:484-493
template <class T> void hard_to_llrs(const Params &c, const uint8_t *input, T *llrs) {
    const T llr = Ops<T>::neg(Ops<T>::one());
    for (int idx = 0; idx < c.n / 8; idx++)
        for (int i = 0; i < 8; i++)
            llrs[idx * 8 + i] = ((input[idx] >> (7 - i)) & 1) ? llr : Ops<T>::neg(llr);
}

// src/decoder.rs:498-509
template <class T> void llrs_to_hard(const Params &c, const T *llrs, uint8_t *output) {
    for (int i = 0; i < c.n / 8; i++) output[i] = 0;
    for (int i = 0; i < c.n; i++)
        if (Ops<T>::hard_bit(llrs[i])) output[i / 8] |= (uint8_t)(1 << (7 - (i % 8)));
}

// One decoder instance per host thread with private scratch, as
// perftest/src/main.rs:39-45 does with rayon workers.
template <class F> void parallel_frames(size_t batch, int nthreads, F &&f) {
    if (nthreads <= 1 || batch <= 1) {
        f(0, batch, 0);
        return;
    }
    std::atomic<size_t> next(0);
    size_t chunk = batch / ((size_t)nthreads * 8);
    if (chunk < 1) chunk = 1;
    if (chunk > 16) chunk = 16;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([&, t]() {
            for (;;) {
                const size_t b0 = next.fetch_add(chunk);
                if (b0 >= batch) break;
                const size_t b1 = b0 + chunk < batch ? b0 + chunk : batch;
                f(b0, b1, t);
            }
        });
    }
    for (auto &x : th) x.join();
}

template <class T>
int decode_ms_batch(int code, const T *llrs, uint8_t *output, size_t batch, size_t maxiters,
                    uint8_t *success, uint32_t *iters, int nthreads) {
    if (!valid(code)) return -1;
    const Params &c = PARAMS[code];
    parallel_frames(batch, nthreads, [&](size_t b0, size_t b1, int) {
        std::vector<T> working(ms_working_len(c));
        std::vector<uint8_t> working_u8(ms_working_u8_len(c));
        for (size_t f = b0; f < b1; f++) {
            size_t it = 0;
            const bool ok = decode_ms<T>(c, llrs + f * c.n, output + f * output_len(c), working.data(),
                                         working_u8.data(), maxiters, &it);
            if (success) success[f] = ok ? 1 : 0;
            if (iters) iters[f] = (uint32_t)it;
        }
    });
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------
// C entry points for ctypes (tests / bench cpu_baseline only).
// ---------------------------------------------------------------------------
extern "C" {

int oracle_code_param(int code, int which) {
    if (!valid(code)) return -1;
    const Params &c = PARAMS[code];
    switch (which) {
        case 0: return c.n;
        case 1: return c.k;
        case 2: return c.p;
        case 3: return c.m;
        case 4: return c.b;
        case 5: return c.edges;
        case 6: return (int)bf_working_len(c);
        case 7: return (int)ms_working_len(c);
        case 8: return (int)ms_working_u8_len(c);
        case 9: return (int)output_len(c);
        default: return -1;
    }
}

// Writes the ordered edge list; returns the edge count.  crc_out (nullable)
// receives the CRC of src/codes/mod.rs:517-535.
int oracle_edges(int code, uint32_t *checks, uint32_t *vars, uint32_t *crc_out) {
    if (!valid(code)) return -1;
    int count = 0;
    uint32_t crc = 0xFFFFFFFFu;
    for_each_edge(PARAMS[code], [&](int, int check, int var) {
        if (checks) checks[count] = (uint32_t)check;
        if (vars) vars[count] = (uint32_t)var;
        crc = crc32_u16(crc, (uint32_t)check);
        crc = crc32_u16(crc, (uint32_t)var);
        count++;
    });
    if (crc_out) *crc_out = crc;
    return count;
}

// word = 8, 32 or 64 selects the u8/u32/u64 implementation (encoder.rs:41/107/189).
int oracle_copy_encode(int code, const uint8_t *data, uint8_t *codeword, int word) {
    if (!valid(code)) return -1;
    const Params &c = PARAMS[code];
    if (!c.gen) return -2;      // no generator: the reference cannot encode this code
    memmove(codeword, data, c.k / 8);
    if (word == 8) {
        encode_parity_u8(c, codeword, codeword + c.k / 8);
    } else if (word == 32) {
        std::vector<uint32_t> par((c.n - c.k) / 32);
        encode_parity_u32(c, codeword, par.data());
        memcpy(codeword + c.k / 8, par.data(), (c.n - c.k) / 8);
    } else if (word == 64) {
        std::vector<uint64_t> par((c.n - c.k) / 64);
        encode_parity_u64(c, codeword, par.data());
        memcpy(codeword + c.k / 8, par.data(), (c.n - c.k) / 8);
    } else {
        return -2;
    }
    return 0;
}

int oracle_copy_encode_batch(int code, const uint8_t *data, uint8_t *codewords, size_t batch,
                             int nthreads) {
    if (!valid(code)) return -1;
    const Params &c = PARAMS[code];
    if (!c.gen) return -2;      // no generator: the reference cannot encode this code
    parallel_frames(batch, nthreads, [&](size_t b0, size_t b1, int) {
        for (size_t f = b0; f < b1; f++) {
            uint8_t *cw = codewords + f * (c.n / 8);
            memmove(cw, data + f * (c.k / 8), c.k / 8);
            std::vector<uint64_t> par((c.n - c.k) / 64);
            encode_parity_u64(c, cw, par.data());
            memcpy(cw + c.k / 8, par.data(), (c.n - c.k) / 8);
        }
    });
    return 0;
}

#define ORACLE_MS(SUFFIX, T)                                                                        \
    int oracle_decode_ms_##SUFFIX(int code, const T *llrs, uint8_t *output, T *working,             \
                                  uint8_t *working_u8, size_t maxiters, size_t *iters_run) {        \
        if (!valid(code)) return -1;                                                                \
        return decode_ms<T>(PARAMS[code], llrs, output, working, working_u8, maxiters, iters_run)   \
                   ? 1 : 0;                                                                         \
    }                                                                                               \
    int oracle_decode_ms_##SUFFIX##_batch(int code, const T *llrs, uint8_t *output, size_t batch,   \
                                          size_t maxiters, uint8_t *success, uint32_t *iters,       \
                                          int nthreads) {                                           \
        return decode_ms_batch<T>(code, llrs, output, batch, maxiters, success, iters, nthreads);   \
    }                                                                                               \
    int oracle_hard_to_llrs_##SUFFIX(int code, const uint8_t *input, T *llrs) {                     \
        if (!valid(code)) return -1;                                                                \
        hard_to_llrs<T>(PARAMS[code], input, llrs);                                                 \
        return 0;                                                                                   \
    }                                                                                               \
    int oracle_llrs_to_hard_##SUFFIX(int code, const T *llrs, uint8_t *output) {                    \
        if (!valid(code)) return -1;                                                                \
        llrs_to_hard<T>(PARAMS[code], llrs, output);                                                \
        return 0;                                                                                   \
    }

ORACLE_MS(i8, int8_t)
ORACLE_MS(i16, int16_t)
ORACLE_MS(i32, int32_t)
ORACLE_MS(f32, float)
ORACLE_MS(f64, double)

int oracle_decode_bf(int code, const uint8_t *input, uint8_t *output, uint8_t *working,
                     size_t maxiters, size_t *iters_run) {
    if (!valid(code)) return -1;
    return decode_bf(PARAMS[code], input, output, working, maxiters, iters_run) ? 1 : 0;
}

int oracle_decode_bf_batch(int code, const uint8_t *input, uint8_t *output, size_t batch,
                           size_t maxiters, uint8_t *success, uint32_t *iters, int nthreads) {
    if (!valid(code)) return -1;
    const Params &c = PARAMS[code];
    parallel_frames(batch, nthreads, [&](size_t b0, size_t b1, int) {
        std::vector<uint8_t> working(bf_working_len(c));
        for (size_t f = b0; f < b1; f++) {
            size_t it = 0;
            const bool ok = decode_bf(c, input + f * (c.n / 8), output + f * output_len(c),
                                      working.data(), maxiters, &it);
            if (success) success[f] = ok ? 1 : 0;
            if (iters) iters[f] = (uint32_t)it;
        }
    });
    return 0;
}

// Exposed for the test mirroring src/decoder.rs:607-645.
int oracle_decode_erasures(int code, uint8_t *codeword, uint8_t *working, size_t maxiters,
                           size_t *iters_run) {
    if (!valid(code)) return -1;
    return decode_erasures(PARAMS[code], codeword, working, maxiters, iters_run) ? 1 : 0;
}

}  // extern "C"
