// One-time host expansion of the compact code descriptions (see code_tables.h).
#include "code_tables.h"

#include <mutex>
#include <utility>

#include "ccsds_tables.h"

namespace ldpc {
namespace {

struct Raw {
    const char *name;
    int n, k, p, m, b, edges;
    const uint8_t *proto;     // [3][4][11]
    const uint16_t *phi;      // [4][26] or nullptr (TC codes have no permutation blocks)
    const uint64_t *gen;
};

// Literal parameters: reference src/codes/mod.rs:109-241.  phi table per M:
// src/codes/mod.rs:469-478; prototype per rate: src/codes/mod.rs:480-486.
const Raw kRaw[kNumCodes] = {
    {"TC128", 128, 64, 0, 16, 16, 512, ccsds_proto_tc128, nullptr, ccsds_gen_tc128},
    {"TC256", 256, 128, 0, 32, 32, 1024, ccsds_proto_tc256, nullptr, ccsds_gen_tc256},
    {"TC512", 512, 256, 0, 64, 64, 2048, ccsds_proto_tc512, nullptr, ccsds_gen_tc512},
    {"TM1280", 1280, 1024, 128, 128, 32, 4992, ccsds_proto_tm_r45, ccsds_phi_m128, ccsds_gen_tm1280},
    {"TM1536", 1536, 1024, 256, 256, 64, 5888, ccsds_proto_tm_r23, ccsds_phi_m256, ccsds_gen_tm1536},
    {"TM2048", 2048, 1024, 512, 512, 128, 7680, ccsds_proto_tm_r12, ccsds_phi_m512, ccsds_gen_tm2048},
    {"TM5120", 5120, 4096, 512, 512, 128, 19968, ccsds_proto_tm_r45, ccsds_phi_m512, ccsds_gen_tm5120},
    {"TM6144", 6144, 4096, 1024, 1024, 256, 23552, ccsds_proto_tm_r23, ccsds_phi_m1024, ccsds_gen_tm6144},
    {"TM8192", 8192, 4096, 2048, 2048, 512, 30720, ccsds_proto_tm_r12, ccsds_phi_m2048, ccsds_gen_tm8192},
    // k = 16384 (CCSDS 131.0-B: M = 2048 / 4096 / 8192 for rates 4/5, 2/3, 1/2; compact_parity_checks.rs:84-96).  The
    // reference ships their prototypes and phi tables but no generators (src/lib.rs:81-83): p = M, b = M/4 and
    // edges = blocks x M follow the pattern of the six smaller TM codes.
    {"TM20480", 20480, 16384, 2048, 2048, 512, 39 * 2048, ccsds_proto_tm_r45, ccsds_phi_m2048, nullptr},
    {"TM24576", 24576, 16384, 4096, 4096, 1024, 23 * 4096, ccsds_proto_tm_r23, ccsds_phi_m4096, nullptr},
    {"TM32768", 32768, 16384, 8192, 8192, 2048, 15 * 8192, ccsds_proto_tm_r12, ccsds_phi_m8192, nullptr},
};

CodeInfo g_info[kNumCodes];
std::once_flag g_once;

// Block enumeration in the reference iterator's order: prototype rows, then
// columns, then the summed sub-prototypes of a cell until its first zero entry
// (reference src/codes/mod.rs:295-361).
void build_all() {
    for (int ci = 0; ci < kNumCodes; ci++) {
        const Raw &r = kRaw[ci];
        CodeInfo &c = g_info[ci];
        c.name = r.name;
        c.n = r.n; c.k = r.k; c.p = r.p; c.m = r.m; c.b = r.b;
        c.edges = r.edges;
        c.checks = r.n + r.p - r.k;
        c.vars = r.n + r.p;
        c.rows = c.checks / c.m;
        c.cols = c.vars / c.m;
        c.gen = r.gen;
        c.n_blocks = 0;
        int offset = 0;
        for (int row = 0; row < 4; row++) {
            for (int col = 0; col < 11; col++) {
                for (int sub = 0; sub < 3; sub++) {
                    const uint8_t e = r.proto[(sub * 4 + row) * 11 + col];
                    if (e == 0) break;
                    const int kind = e & CCSDS_KIND_MASK;
                    const int val = e & CCSDS_VAL_MASK;
                    Block &b = c.blocks[c.n_blocks++];
                    b.row = row; b.col = col;
                    b.shift = 0; b.theta = 0;
                    b.phi[0] = b.phi[1] = b.phi[2] = b.phi[3] = 0;
                    b.edge_offset = offset;
                    if (kind == CCSDS_KIND_IDENT) {
                        b.kind = kIdentity;
                        b.shift = val % c.m;
                    } else {
                        b.kind = kPermutation;
                        b.theta = ccsds_theta_k[val];
                        for (int j = 0; j < 4; j++) b.phi[j] = r.phi[j * 26 + val] % (c.m / 4);
                    }
                    offset += c.m;
                }
            }
        }
        // degrees
        std::vector<int> vd(c.vars, 0), cd(c.checks, 0);
        std::vector<uint32_t> chk, var;
        expand_edges(c, chk, var);
        for (size_t i = 0; i < chk.size(); i++) { vd[var[i]]++; cd[chk[i]]++; }
        c.max_var_degree = 0; c.max_check_degree = 0;
        for (int d : vd) if (d > c.max_var_degree) c.max_var_degree = d;
        for (int d : cd) if (d > c.max_check_degree) c.max_check_degree = d;
    }
}

inline int block_pi(const CodeInfo &c, const Block &b, int i) {
    if (b.kind == kIdentity) return (i + b.shift) % c.m;
    const int q = c.m / 4;
    const int j = i / q;
    return q * ((b.theta + j) % 4) + ((b.phi[j] + i) % q);
}

}  // namespace

const CodeInfo *code_info(int code) {
    if (code < 0 || code >= kNumCodes) return nullptr;
    std::call_once(g_once, build_all);
    return &g_info[code];
}

void expand_edges(const CodeInfo &c, std::vector<uint32_t> &check, std::vector<uint32_t> &var) {
    check.clear(); var.clear();
    check.reserve(c.edges); var.reserve(c.edges);
    for (int bi = 0; bi < c.n_blocks; bi++) {
        const Block &b = c.blocks[bi];
        for (int i = 0; i < c.m; i++) {
            check.push_back((uint32_t)(b.row * c.m + i));
            var.push_back((uint32_t)(b.col * c.m + block_pi(c, b, i)));
        }
    }
}

uint32_t edge_crc(const CodeInfo &c) {
    std::vector<uint32_t> chk, var;
    expand_edges(c, chk, var);
    uint32_t crc = 0xFFFFFFFFu;
    auto step = [&crc](uint32_t data) {
        crc ^= data;
        for (int i = 0; i < 16; i++) crc = (crc >> 1) ^ ((crc & 1) ? 0xEDB88320u : 0u);
    };
    for (size_t i = 0; i < chk.size(); i++) { step(chk[i]); step(var[i]); }
    return crc;
}

void build_ell_tables(const CodeInfo &c, std::vector<uint64_t> &var_tab, std::vector<uint64_t> &chk_tab) {
    std::vector<uint32_t> chk, var;
    expand_edges(c, chk, var);
    var_tab.assign((size_t)c.max_var_degree * c.vars, kNoEdge64);
    chk_tab.assign((size_t)c.max_check_degree * c.checks, kNoEdge64);
    std::vector<int> vfill(c.vars, 0), cfill(c.checks, 0);
    for (uint32_t idx = 0; idx < chk.size(); idx++) {
        const uint32_t a = var[idx], ch = chk[idx];
        var_tab[(size_t)vfill[a]++ * c.vars + a] = (uint64_t)idx | ((uint64_t)ch << 32);
        chk_tab[(size_t)cfill[ch]++ * c.checks + ch] = (uint64_t)idx | ((uint64_t)a << 32);
    }
}

namespace {

// Arithmetic in R = GF(2)[x] / (x^Q - 1), Q a power of two >= 32: a Q x Q circulant with first column a is the
// polynomial sum a_z x^z, products of circulants are products in R.  Since x^Q - 1 = (x + 1)^Q, R is a local ring: a is
// a unit iff its weight is odd, and then a^Q = a(x^Q) = a(1) = 1, so a^-1 = a^(Q-1) = a * a^2 * a^4 * ... * a^(Q/2).
struct Poly {
    int q;
    std::vector<uint64_t> w;
    explicit Poly(int q_) : q(q_), w(q_ >= 64 ? q_ / 64 : 1, 0) {}
    bool bit(int z) const { return (w[z >> 6] >> (z & 63)) & 1; }
    void flip(int z) { w[z >> 6] ^= 1ull << (z & 63); }
    bool zero() const { for (uint64_t x : w) if (x) return false; return true; }
    bool unit() const { int p = 0; for (uint64_t x : w) p ^= __builtin_parityll(x); return p != 0; }
    void add(const Poly &o) { for (size_t i = 0; i < w.size(); i++) w[i] ^= o.w[i]; }
    // this += o * x^z
    void add_rot(const Poly &o, int z) {
        if (q < 64) {
            const uint64_t m = (1ull << q) - 1, v = o.w[0];
            w[0] ^= z ? (((v << z) | (v >> (q - z))) & m) : v;
            return;
        }
        const int W = (int)w.size(), ws = z >> 6, bs = z & 63;
        for (int i = 0; i < W; i++) {
            const uint64_t lo = o.w[(i - ws + W) % W], hi = o.w[(i - ws - 1 + 2 * W) % W];
            w[i] ^= bs ? ((lo << bs) | (hi >> (64 - bs))) : lo;
        }
    }
    Poly mul(const Poly &o) const {
        Poly r(q);
        for (int z = 0; z < q; z++) if (bit(z)) r.add_rot(o, z);
        return r;
    }
    Poly square() const {             // a(x)^2 = a(x^2): x^z -> x^(2z mod Q); z and z + Q/2 collide and cancel
        Poly r(q);
        for (int z = 0; z < q; z++) if (bit(z)) r.flip((2 * z) % q);
        return r;
    }
    Poly inverse() const {            // requires unit()
        Poly r(*this), s(*this);
        for (int e = 2; e < q; e <<= 1) { s = s.square(); r = r.mul(s); }
        return r;
    }
};

}  // namespace

bool tm_encoder_table(int code, std::vector<uint32_t> &out) {
    out.clear();
    const CodeInfo *ci = code_info(code);
    if (!ci || ci->p == 0 || ci->rows != 3 || ci->m % 128 != 0 || ci->n - ci->k != 2 * ci->m) return false;
    const CodeInfo &c = *ci;
    const int M = c.m, Q = M / 4, QW = Q / 32;
    const int CA = c.cols - 3, CB = c.cols - 2, CC = c.cols - 1;
    // structure check: which blocks sit in the three parity columns
    std::vector<const Block *> s_blocks, g_blocks;     // S = row 2 / col CB,  G = row 1 / col CC
    int ident_ok = 0;
    for (int bi = 0; bi < c.n_blocks; bi++) {
        const Block &b = c.blocks[bi];
        const bool ident = b.kind == kIdentity && b.shift == 0;
        if (b.col == CA) { if (b.row != 0 || !ident) return false; ident_ok |= 1; }
        else if (b.col == CB) {
            if (b.row == 1) { if (!ident) return false; ident_ok |= 2; }
            else if (b.row == 2) s_blocks.push_back(&b);
            else return false;
        } else if (b.col == CC) {
            if (b.row == 1) g_blocks.push_back(&b);
            else if (b.row == 2) { if (!ident) return false; ident_ok |= 4; }
        } else if (b.row == 0) return false;            // row 0 has no data terms
    }
    if (ident_ok != 7 || s_blocks.empty() || g_blocks.empty()) return false;
    // Every pi_k is a 4 x 4 array of circulants with one monomial per block row: check quarter j, offset i' goes to variable
    // quarter (theta + j) mod 4, offset (phi_j + i') mod Q, i.e. block (j, (theta + j) mod 4) has the first column
    // x^(-phi_j).  A = I + (sum of the S blocks)(sum of the G blocks) is inverted by Gauss-Jordan over the ring
    // (pivots must be units; 4 x 4, so even M = 8192 takes milliseconds where the bit matrix would take seconds).
    auto block_matrix = [&](const std::vector<const Block *> &blocks) {
        std::vector<Poly> m(16, Poly(Q));
        for (const Block *b : blocks)
            for (int j = 0; j < 4; j++) m[(size_t)j * 4 + (b->theta + j) % 4].flip((Q - b->phi[j] % Q) % Q);
        return m;
    };
    const std::vector<Poly> sm = block_matrix(s_blocks), gm = block_matrix(g_blocks);
    std::vector<Poly> a(32, Poly(Q));                     // [row][col 0..3 = A, 4..7 = identity]
    for (int i = 0; i < 4; i++) {
        a[(size_t)i * 8 + i].flip(0);
        a[(size_t)i * 8 + 4 + i].flip(0);
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++)
                if (!sm[(size_t)i * 4 + k].zero() && !gm[(size_t)k * 4 + j].zero())
                    a[(size_t)i * 8 + j].add(sm[(size_t)i * 4 + k].mul(gm[(size_t)k * 4 + j]));
    }
    for (int col = 0; col < 4; col++) {
        int piv = -1;
        for (int r = col; r < 4; r++) if (a[(size_t)r * 8 + col].unit()) { piv = r; break; }
        if (piv < 0) return false;
        if (piv != col) for (int j = 0; j < 8; j++) std::swap(a[(size_t)piv * 8 + j], a[(size_t)col * 8 + j]);
        const Poly inv = a[(size_t)col * 8 + col].inverse();
        for (int j = 0; j < 8; j++) a[(size_t)col * 8 + j] = inv.mul(a[(size_t)col * 8 + j]);
        for (int r = 0; r < 4; r++) {
            if (r == col || a[(size_t)r * 8 + col].zero()) continue;
            const Poly f = a[(size_t)r * 8 + col];
            for (int j = 0; j < 8; j++) a[(size_t)r * 8 + j].add(f.mul(a[(size_t)col * 8 + j]));
        }
    }
    out.assign((size_t)16 * QW, 0);
    for (int qi = 0; qi < 4; qi++)
        for (int qj = 0; qj < 4; qj++)
            for (int z = 0; z < Q; z++)
                if (a[(size_t)qi * 8 + 4 + qj].bit(z)) out[(size_t)(qi * 4 + qj) * QW + (z >> 5)] |= 1u << (z & 31);
    return true;
}

bool tm_encoder_lut(int code, std::vector<uint32_t> &lut) {
    lut.clear();
    std::vector<uint32_t> col;
    if (!tm_encoder_table(code, col)) return false;
    const int M = code_info(code)->m, Q = M / 4, QW = Q / 32, MW = M / 32;
    lut.assign((size_t)512 * MW, 0);
    auto row = [&](int v, int qj, int nib) { return &lut[((size_t)v * 32 + qj * 8 + nib) * MW]; };
    // single-bit rows first (v = 1 << e), the rest are XORs of those
    for (int qj = 0; qj < 4; qj++)
        for (int nib = 0; nib < 8; nib++) {
            for (int e = 0; e < 4; e++)
                for (int qi = 0; qi < 4; qi++)
                    for (int x = 0; x < Q; x++) {
                        const int z = ((x - 4 * nib - e) % Q + Q) % Q;
                        if ((col[(size_t)(qi * 4 + qj) * QW + (z >> 5)] >> (z & 31)) & 1)
                            row(1 << e, qj, nib)[qi * QW + (x >> 5)] |= 1u << (x & 31);
                    }
            for (int v = 3; v < 16; v++) {
                if ((v & (v - 1)) == 0) continue;
                const int low = v & -v;
                for (int w = 0; w < MW; w++) row(v, qj, nib)[w] = row(v ^ low, qj, nib)[w] ^ row(low, qj, nib)[w];
            }
        }
    return true;
}

bool tc_encoder_lut(int code, int group_bits, std::vector<uint32_t> &lut) {
    lut.clear();
    const CodeInfo *ci = code_info(code);
    if (!ci || ci->p != 0 || (group_bits != 4 && group_bits != 8)) return false;
    const CodeInfo &c = *ci;
    const int r = c.n - c.k, PW = r / 32, W64 = r / 64, nv = 1 << group_bits, groups = c.k / group_bits;
    // parity row of one data bit, as memory-order words
    std::vector<uint32_t> rows((size_t)c.k * PW, 0);
    for (int i = 0; i < c.k; i++) {
        const int crow = i / c.b, o = i % c.b;
        for (int j = 0; j < r; j++) {
            const int src = (j / c.b) * c.b + ((j % c.b) - o + c.b) % c.b;
            if ((c.gen[(size_t)crow * W64 + src / 64] >> (63 - src % 64)) & 1)
                rows[(size_t)i * PW + j / 32] |= 1u << (8 * ((j / 8) % 4) + 7 - j % 8);
        }
    }
    lut.assign((size_t)groups * nv * PW, 0);
    for (int g = 0; g < groups; g++)
        for (int v = 1; v < nv; v++) {
            const int low = v & -v, e = __builtin_ctz(low);
            // bit e of the group value is data bit (first bit of the group) + group_bits - 1 - e
            const int bit = g * group_bits + group_bits - 1 - e;
            for (int w = 0; w < PW; w++)
                lut[((size_t)g * nv + v) * PW + w] = lut[((size_t)g * nv + (v ^ low)) * PW + w] ^ rows[(size_t)bit * PW + w];
        }
    return true;
}

bool tc_rot_encoder_lut(int code, std::vector<uint32_t> &lut) {
    lut.clear();
    const CodeInfo *ci = code_info(code);
    if (!ci || ci->p != 0 || ci->b % 32 != 0 || ci->k != 4 * ci->b) return false;
    std::vector<uint32_t> full;
    if (!tc_encoder_lut(code, 8, full)) return false;
    const int KW = (ci->n - ci->k) / 32, BB = ci->b / 8;
    lut.assign((size_t)4 * 256 * KW, 0);
    for (int crow = 0; crow < 4; crow++)
        for (int v = 0; v < 256; v++)
            for (int w = 0; w < KW; w++) lut[((size_t)crow * 256 + v) * KW + w] = full[((size_t)(crow * BB) * 256 + v) * KW + w];
    return true;
}

void host_encode_generator(int code, const uint8_t *data, uint8_t *parity) {
    const CodeInfo &c = *code_info(code);
    const int r = c.n - c.k, W64 = r / 64;
    for (int i = 0; i < r / 8; i++) parity[i] = 0;
    for (int i = 0; i < c.k; i++) {
        if (!((data[i / 8] >> (7 - i % 8)) & 1)) continue;
        const int crow = i / c.b, o = i % c.b;
        for (int j = 0; j < r; j++) {
            const int src = (j / c.b) * c.b + ((j % c.b) - o + c.b) % c.b;      // row crow, blocks rotated right by o
            if ((c.gen[(size_t)crow * W64 + src / 64] >> (63 - src % 64)) & 1) parity[j / 8] ^= (uint8_t)(0x80 >> (j % 8));
        }
    }
}

namespace {
uint32_t brev32(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++)
        if ((x >> i) & 1) r |= 1u << (31 - i);
    return r;
}
uint32_t funnel_r(uint32_t lo, uint32_t hi, unsigned s) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
uint32_t funnel_l(uint32_t lo, uint32_t hi, unsigned s) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
}  // namespace

bool host_encode_tables(int code, const uint8_t *data, uint8_t *parity) {
    const CodeInfo *ci = code_info(code);
    if (!ci) return false;
    const CodeInfo &c = *ci;
    if (code == 1 || code == 2) {                            // TC256 / TC512: byte rows of block position 0, rotated by whole bytes
        std::vector<uint32_t> lut;
        if (!tc_rot_encoder_lut(code, lut)) return false;
        const int PB = (c.n - c.k) / 8, BB = c.b / 8;        // parity bytes, bytes per circulant block
        std::vector<uint8_t> out(PB, 0);
        for (int j = 0; j < c.k / 8; j++) {
            const int crow = j / BB, y = j % BB;
            const uint8_t *row = reinterpret_cast<const uint8_t *>(&lut[((size_t)crow * 256 + data[j]) * (PB / 4)]);   // little-endian words = memory order
            for (int blk = 0; blk < PB / BB; blk++)
                for (int q = 0; q < BB; q++) out[blk * BB + q] ^= row[blk * BB + ((q - y) & (BB - 1))];
        }
        for (int i = 0; i < PB; i++) parity[i] = out[i];
        return true;
    }
    if (c.p == 0) {                                          // TC codes: encode_tc_lut_kernel
        const int GB = tc_encoder_group_bits(code), NV = 1 << GB, KW = c.k / 32;
        std::vector<uint32_t> lut;
        if (!tc_encoder_lut(code, GB, lut)) return false;
        std::vector<uint32_t> p(KW, 0);
        for (int j = 0; j < c.k / 8; j++)
            for (int w = 0; w < KW; w++) {
                if (GB == 8) p[w] ^= lut[((size_t)j * NV + data[j]) * KW + w];
                else p[w] ^= lut[((size_t)(2 * j) * NV + (data[j] >> 4)) * KW + w] ^ lut[((size_t)(2 * j + 1) * NV + (data[j] & 15)) * KW + w];
            }
        for (int i = 0; i < (c.n - c.k) / 8; i++) parity[i] = (uint8_t)(p[i / 4] >> (8 * (i % 4)));
        return true;
    }
    // TM codes: encode_tm_kernel (both forms of the dense product)
    std::vector<uint32_t> ainv, lut;
    if (!tm_encoder_table(code, ainv) || !tm_encoder_lut(code, lut)) return false;
    const int M = c.m, Q = M / 4, QW = Q / 32, MW = M / 32, KC = c.cols - 3, CB = c.cols - 2, CC = c.cols - 1;
    auto slot = [&](int col) { return col < KC ? col : (col == CB ? KC : KC + 1); };
    std::vector<uint32_t> lo((size_t)(KC + 2) * MW), hi((size_t)(KC + 2) * MW);
    auto store = [&](int base, int w, uint32_t v) { lo[base + w] = v; hi[base + ((w & ~(QW - 1)) | ((w - 1) & (QW - 1)))] = v; };
    auto window = [&](const Block &b, int w) {
        const int e0 = w * 32, qa = e0 / Q, off = e0 % Q;
        const int qv = (b.theta + qa) & 3, s0 = (b.phi[qa] + off) & (Q - 1);
        const int i = slot(b.col) * MW + qv * QW + (s0 >> 5);
        return funnel_r(lo[i], hi[i], s0 & 31);
    };
    std::vector<uint32_t> dw((size_t)KC * MW), t1(MW), sv(MW), pc(MW, 0), pc2(MW, 0);
    for (int col = 0; col < KC; col++)
        for (int w = 0; w < MW; w++) {
            const uint8_t *p = &data[4 * (col * MW + w)];
            dw[col * MW + w] = brev32(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]);
            store(col * MW, w, dw[col * MW + w]);
        }
    for (int w = 0; w < MW; w++) {
        uint32_t x1 = 0, x2 = 0;
        for (int bi = 0; bi < c.n_blocks; bi++) {
            const Block &b = c.blocks[bi];
            if (b.col >= KC) continue;
            const uint32_t v = b.kind == kPermutation ? window(b, w) : dw[b.col * MW + w];
            if (b.row == 1) x1 ^= v; else x2 ^= v;
        }
        t1[w] = x1; sv[w] = x2;
    }
    for (int w = 0; w < MW; w++) store(KC * MW, w, t1[w]);
    for (int w = 0; w < MW; w++)
        for (int bi = 0; bi < c.n_blocks; bi++)
            if (c.blocks[bi].col == CB && c.blocks[bi].row == 2) sv[w] ^= window(c.blocks[bi], w);
    for (int j = 0; j < MW; j++) {
        const int qj = j / QW, wq = j % QW;
        for (int w = 0; w < MW; w++) {
            const int qi = w / QW, wx = w % QW, w0 = (wx - wq) & (QW - 1);
            const uint32_t *col = &ainv[(size_t)(qi * 4 + qj) * QW];
            for (uint32_t D = sv[j]; D; D &= D - 1) pc[w] ^= funnel_l(col[(w0 - 1) & (QW - 1)], col[w0], __builtin_ctz(D));
            for (int nib = 0; nib < 8; nib++)
                pc2[w] ^= lut[((size_t)((sv[j] >> (4 * nib)) & 15) * 32 + qj * 8 + nib) * MW + ((w & ~(QW - 1)) | ((w - wq) & (QW - 1)))];
        }
    }
    if (pc != pc2) return false;
    for (int w = 0; w < MW; w++) store((KC + 1) * MW, w, pc[w]);
    for (int w = 0; w < MW; w++) {
        uint32_t pa = 0, pb = t1[w];
        for (int bi = 0; bi < c.n_blocks; bi++) {
            const Block &b = c.blocks[bi];
            if (b.col == CC && b.row == 1) pb ^= window(b, w);
            if (b.col == CC && b.row == 0) pa ^= b.kind == kPermutation ? window(b, w) : pc[w];
        }
        const uint32_t ra = brev32(pa), rb = brev32(pb);
        for (int k = 0; k < 4; k++) {
            parity[4 * w + k] = (uint8_t)(ra >> (24 - 8 * k));
            parity[M / 8 + 4 * w + k] = (uint8_t)(rb >> (24 - 8 * k));
        }
    }
    return true;
}

}  // namespace ldpc
