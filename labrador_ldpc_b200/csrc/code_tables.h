// Host-side description of the CCSDS codes -- the reference's nine plus the three k = 16384 TM codes whose parity-check
// constants the reference carries without supporting them (src/lib.rs:81-83) -- and the one-time expansion of
// the compact parity-check prototypes into block descriptors and edge tables.
//
// Replaces (as data, built once per process instead of re-derived per edge):
//   CodeParams / *_PARAMS            reference src/codes/mod.rs:69-241
//   ParityIter::next + iterator setup reference src/codes/mod.rs:275-362, 444-494
//   compact_generator()              reference src/codes/mod.rs:412-424
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ldpc {

constexpr int kNumCodes = 12;       // 0..8: enum LDPCCode of the reference (src/codes/mod.rs:37-66); 9..11: TM20480 / TM24576 / TM32768
constexpr int kNumRefCodes = 9;
constexpr int kMaxBlocks = 40;     // r4/5 prototype has 39 non-zero blocks
constexpr uint32_t kNoEdge = 0xFFFFFFFFu;

enum BlockKind : int { kIdentity = 1, kPermutation = 2 };

// One non-zero MxM block of H.  Edges of the block are i = 0..M-1:
//   check = row*M + i,  idx = edge_offset + i,
//   var   = col*M + pi(i)  with
//   identity:    pi(i) = (i + shift) mod M
//   permutation: pi(i) = Q*((theta + i/Q) mod 4) + ((phi[i/Q] + i) mod Q),  Q = M/4
struct Block {
    int row, col, kind;
    int shift;          // identity only
    int theta;          // permutation only
    int phi[4];         // permutation only
    int edge_offset;    // idx of the block's first edge in reference order
};

struct CodeInfo {
    const char *name;
    int n, k, p, m, b;
    int edges;               // paritycheck_sum
    int checks;              // n + p - k
    int vars;                // n + p
    int rows, cols;          // active prototype rows / columns
    const uint64_t *gen;     // compact generator, (k/b) x ((n-k)/64) words, MSB = parity bit 0; nullptr for the k = 16384 codes
    int n_blocks;
    Block blocks[kMaxBlocks];
    int max_var_degree, max_check_degree;

    // reference src/decoder.rs:93-116
    size_t bf_working_len() const { return (size_t)n + p; }
    size_t ms_working_len() const { return 2 * (size_t)edges + 3 * (size_t)n + 3 * (size_t)p - 2 * (size_t)k; }
    size_t ms_working_u8_len() const { return (size_t)(n + p - k) / 8; }
    size_t output_len() const { return (size_t)(n + p) / 8; }
};

// Returns nullptr for an out-of-range code.
const CodeInfo *code_info(int code);

// Ordered edge list (reference order): check[idx], var[idx].
void expand_edges(const CodeInfo &c, std::vector<uint32_t> &check, std::vector<uint32_t> &var);

// CRC-32 over the ordered edge list as in reference src/codes/mod.rs:508-535.
uint32_t edge_crc(const CodeInfo &c);

// ELL tables for the generic kernels (64-bit entries: the k = 16384 codes have more than 2^16 edges).
//   var_tab[j*vars + a]   = idx | check << 32   j-th edge of variable a in ascending idx order
//   chk_tab[j*checks + c] = idx | var   << 32   j-th edge of check c in ascending idx order
// Missing entries are kNoEdge64.
constexpr uint64_t kNoEdge64 = ~0ull;
void build_ell_tables(const CodeInfo &c, std::vector<uint64_t> &var_tab, std::vector<uint64_t> &chk_tab);

// Table of the parity-check-based TM encoder (encode_tm.cu).  In every TM prototype the three parity block
// columns (CA, CB transmitted, CC punctured) sit in H as
//     row 0:  I(CA)                +  (I + P0)(CC)
//     row 1:  [data]  +  I(CB)     +  (P1 + P2 + P3)(CC)
//     row 2:  [data]  +  S(CB)     +  I(CC)              S = P6 + P7
// so with t1 / t2 = the data terms of rows 1 / 2:  A * p_CC = t2 + S t1,  A = I + S (P1 + P2 + P3), and then
// p_CB = t1 + (P1 + P2 + P3) p_CC, p_CA = (I + P0) p_CC.  Every pi_k maps quarter to quarter by a rotation, so A^-1
// is a 4 x 4 array of Q x Q circulants (Q = M/4) and is fully described by the first column of each:
//     out[(qi * 4 + qj) * (Q / 32) + w], bit t  =  A^-1 [qi * Q + 32 w + t] [qj * Q]
// (Q = 32 for M = 128: one word).  Returns false (out empty) if the code is not a TM code with this structure or
// A is singular.  The result is the same systematic codeword the compact generator produces (the parity part of
// H is invertible, so the parity bits of a data word are unique); tests compare the two bit for bit.
bool tm_encoder_table(int code, std::vector<uint32_t> &out);

// The same A^-1 as a nibble lookup table ("four Russians"): for every source quarter qj, nibble position
// nib (0..7) inside a 32-bit word of s and nibble value v, the XOR of the (up to four) rotated columns those
// bits select, for word offset 0:
//     lut[((v * 4 + qj) * 8 + nib) * (M / 32) + qi * (Q / 32) + w], bit t
//         = XOR over bits e of v of  A^-1 [qi * Q + (32 w + t - 4 nib - e) mod Q] [qj * Q]
// A word of s at in-quarter word index wq then contributes lut[..][(w - wq) mod (Q/32)] to word w of quarter qi
// of the product.  512 rows of M bits: 8 KB (M = 128) ... 128 KB (M = 2048).
bool tm_encoder_lut(int code, std::vector<uint32_t> &lut);

// Lookup table of the TC-code encoder (encode.cu: encode_tc_lut_kernel): the parity contribution of every value of
// every group of `group_bits` (4 or 8) consecutive data bits, i.e. the XOR of the generator rows those bits select
// (reference src/encoder.rs:42-82: the row of data bit crow*b + o is compact row crow with every b-bit block rotated
// right by o).  Data bytes are taken as they lie in memory (MSB = lowest bit index); a row is (n-k)/32 words holding
// the parity bytes in memory order (little-endian words), so it can be XORed and stored without any byte swapping:
//     lut[(g * 2^group_bits + v) * (n-k)/32 + w],   g = group index (byte j, or nibble 2j = high / 2j+1 = low half)
bool tc_encoder_lut(int code, int group_bits, std::vector<uint32_t> &lut);
// group size the table kernel is instantiated with: nibble rows for TC128 (16 rows = one sweep of the 32 banks, so a
// load never conflicts) and TC512 (table size), byte rows for TC256
inline int tc_encoder_group_bits(int code) { return code == 1 ? 8 : 4; }
// TC256 / TC512 (b = 32 / 64: a circulant block is 4 / 8 whole bytes): a data byte at byte position y of its block row
// contributes what the byte at position 0 contributes with every parity block rotated by y bytes, so only the rows of
// position 0 are tabulated -- 4 block rows x 256 values x (n-k)/8 bytes = 16 / 32 KB:
//     lut[(crow * 256 + v) * (n-k)/32 + w]
bool tc_rot_encoder_lut(int code, std::vector<uint32_t> &lut);

// Host-side models of the encoders, for the CPU tests (tests/test_capi_host.py): the parity bytes of one data block
//   * host_encode_generator: by the compact generator, the reference's algorithm (src/encoder.rs:42-82);
//   * host_encode_tables:    by the derived tables, word for word the way the kernels use them -- TM codes: windows,
//     A^-1 as first columns (funnel shifts) AND as nibble lookup table, which must agree; TC codes: the byte / nibble
//     table.  Returns false if a table is missing or the two TM forms disagree.
void host_encode_generator(int code, const uint8_t *data, uint8_t *parity);
bool host_encode_tables(int code, const uint8_t *data, uint8_t *parity);

}  // namespace ldpc
