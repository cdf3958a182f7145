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

void build_ell_tables(const CodeInfo &c, std::vector<uint32_t> &var_tab, std::vector<uint32_t> &chk_tab) {
    std::vector<uint32_t> chk, var;
    expand_edges(c, chk, var);
    var_tab.assign((size_t)c.max_var_degree * c.vars, kNoEdge);
    chk_tab.assign((size_t)c.max_check_degree * c.checks, kNoEdge);
    std::vector<int> vfill(c.vars, 0), cfill(c.checks, 0);
    for (uint32_t idx = 0; idx < chk.size(); idx++) {
        const uint32_t a = var[idx], ch = chk[idx];
        var_tab[(size_t)vfill[a]++ * c.vars + a] = idx | (ch << 16);
        chk_tab[(size_t)cfill[ch]++ * c.checks + ch] = idx | (a << 16);
    }
}

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
    // A = I + S G as bit rows, augmented with the four right-hand sides e_{qj Q}
    const int RW = M / 64 + 1;
    std::vector<uint64_t> a((size_t)M * RW, 0);
    auto flip = [&](int r, int col) { a[(size_t)r * RW + (col >> 6)] ^= 1ull << (col & 63); };
    for (int i = 0; i < M; i++) {
        flip(i, i);
        for (const Block *s : s_blocks)
            for (const Block *g : g_blocks) flip(i, block_pi(c, *g, block_pi(c, *s, i)));
    }
    for (int qj = 0; qj < 4; qj++) flip(qj * Q, M + qj);
    // Gauss-Jordan over GF(2)
    for (int col = 0; col < M; col++) {
        int piv = -1;
        for (int r = col; r < M; r++)
            if ((a[(size_t)r * RW + (col >> 6)] >> (col & 63)) & 1) { piv = r; break; }
        if (piv < 0) return false;
        if (piv != col)
            for (int w = 0; w < RW; w++) std::swap(a[(size_t)piv * RW + w], a[(size_t)col * RW + w]);
        const uint64_t *prow = &a[(size_t)col * RW];
        const int w0 = col >> 6;
        for (int r = 0; r < M; r++) {
            if (r == col) continue;
            uint64_t *row = &a[(size_t)r * RW];
            if ((row[w0] >> (col & 63)) & 1)
                for (int w = w0; w < RW; w++) row[w] ^= prow[w];
        }
    }
    out.assign((size_t)16 * QW, 0);
    for (int qi = 0; qi < 4; qi++)
        for (int qj = 0; qj < 4; qj++)
            for (int z = 0; z < Q; z++) {
                const int bit = M + qj;
                if ((a[(size_t)(qi * Q + z) * RW + (bit >> 6)] >> (bit & 63)) & 1)
                    out[(size_t)(qi * 4 + qj) * QW + (z >> 5)] |= 1u << (z & 31);
            }
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
