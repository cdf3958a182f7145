"""CPU-side checks of the product library: it loads, exports every symbol the header
declares, and its host logic (size getters, table expansion, argument validation) is right.
No compute call succeeds here (no GPU in this container)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["TC128", "TC256", "TC512", "TM1280", "TM1536", "TM2048", "TM5120", "TM6144", "TM8192"]


def header_symbols():
    text = open(os.path.join(ROOT, "include", "labrador_ldpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(labrador_ldpc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ldpc):
    syms = header_symbols()
    assert len(syms) >= 21 + 20
    for s in syms:
        assert hasattr(ldpc.lib, s), s
    # the reference's 21 entry points (capi/src/lib.rs:15-179)
    ref21 = ["code_n", "code_k", "encode", "copy_encode", "bf_working_len", "ms_working_len",
             "ms_working_u8_len", "output_len", "decode_bf"]
    for t in ("i8", "i16", "f32", "f64"):
        ref21 += ["decode_ms_" + t, "hard_to_llrs_" + t, "llrs_to_hard_" + t]
    assert len(ref21) == 21
    for s in ref21:
        assert "labrador_ldpc_" + s in syms


def test_size_getters_match_reference_table(ldpc, kats):
    for code, name in enumerate(NAMES):
        P = kats["params"][name]
        c = ldpc.LDPCCode(code)
        assert c.n() == P["n"] and c.k() == P["k"]
        assert c.punctured_bits() == P["punctured_bits"]
        assert c.paritycheck_sum() == P["paritycheck_sum"]
        assert c.decode_bf_working_len() == P["decode_bf_working_len"]
        assert c.decode_ms_working_len() == P["decode_ms_working_len"]
        assert c.decode_ms_working_u8_len() == P["decode_ms_working_u8_len"]
        assert c.output_len() == P["output_len"]
    assert ldpc.lib.labrador_ldpc_code_n(12) == 0 and ldpc.lib.labrador_ldpc_code_n(-1) == 0


def test_edge_table_crc_goldens(ldpc, kats):
    # the device edge tables are expanded from the same block list; its CRC must hit
    # the reference's goldens (src/codes/mod.rs:521-523)
    for code in range(9):
        assert ldpc.LDPCCode(code).edge_table_crc() == kats["edge_crc32"][code]


def test_encoder_tables_against_generator_and_oracle(ldpc, kats):
    """The derived encoder tables (TM: block-circulant inverse of the parity part of H as first columns and as
    nibble lookup table; TC: byte / nibble table) reproduce the generator encoder and the oracle on the host,
    for the reference's known answers (src/encoder.rs:361-527) and random data."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    oracle = pyoracle.Oracle()
    L = ldpc.lib
    L.labrador_ldpc_host_encode_model.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.labrador_ldpc_host_encode_model.restype = ctypes.c_int
    names = ["TC128", "TC256", "TC512", "TM1280", "TM1536", "TM2048", "TM5120", "TM6144", "TM8192"]
    rng = np.random.default_rng(7)
    for code in range(9):
        c = ldpc.LDPCCode(code)
        kb, pb = c.k() // 8, (c.n() - c.k()) // 8
        cases = [(np.arange(kb) % 256).astype(np.uint8), np.zeros(kb, np.uint8), np.full(kb, 0xFF, np.uint8)]
        cases += [rng.integers(0, 256, kb, dtype=np.uint8) for _ in range(3)]
        one = np.zeros(kb, np.uint8)
        one[kb - 1] = 1                      # a single data bit (the last one)
        cases.append(one)
        for i, data in enumerate(cases):
            pt, pg = np.zeros(pb, np.uint8), np.zeros(pb, np.uint8)
            assert L.labrador_ldpc_host_encode_model(code, data.ctypes.data, pt.ctypes.data, pg.ctypes.data) == 0
            want = oracle.copy_encode(code, data)[kb:]
            assert np.array_equal(pg, want), (names[code], i, "generator model")
            assert np.array_equal(pt, want), (names[code], i, "table model")
            if i == 0:
                assert pt.tolist() == kats["encode_parity"][names[code]]
    assert L.labrador_ldpc_host_encode_model(12, cases[0].ctypes.data, pt.ctypes.data, pg.ctypes.data) < 0


def test_k16384_codes_tables_and_sparse_encoder(ldpc):
    """The three k = 16384 codes (extension; the reference has their parity-check constants but no parameters, no
    generators and therefore NO golden of any kind).  What can be pinned on the host: the oracle and the product expand
    the same edge list (CRC), the degrees follow the prototypes, and the product's sparse-H encoder tables produce
    codewords that satisfy every parity check of the ORACLE's edge list -- H c = 0 with the data bits in place is the
    definition of the systematic encoder, and the parity part of H is invertible, so the answer is unique."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    oracle = pyoracle.Oracle()
    L = ldpc.lib
    L.labrador_ldpc_host_encode_model.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.labrador_ldpc_host_encode_model.restype = ctypes.c_int
    rng = np.random.default_rng(11)
    expect = {9: (20480, 16384, 2048, 2048, 39), 10: (24576, 16384, 4096, 4096, 23), 11: (32768, 16384, 8192, 8192, 15)}
    for code, (n, k, p, m, blocks) in expect.items():
        c = ldpc.LDPCCode(code)
        assert (c.n(), c.k(), c.punctured_bits(), c.paritycheck_sum()) == (n, k, p, blocks * m)
        assert (oracle.n(code), oracle.k(code), oracle.p(code), oracle.m(code)) == (n, k, p, m)
        assert c.output_len() == (n + p) // 8 and c.decode_ms_working_len() == 2 * blocks * m + 3 * n + 3 * p - 2 * k
        cnt, checks, vars_, crc = oracle.edges(code)
        assert cnt == blocks * m and crc == c.edge_table_crc()
        # check degrees 3 (row 0) and 2 * (blocks - 3) / 2 ... as in the smaller code of the same rate
        deg_c = np.bincount(checks, minlength=n + p - k)
        small = {9: 6, 10: 7, 11: 8}[code]
        _, ch_s, _, _ = oracle.edges(small)
        assert sorted(set(deg_c.tolist())) == sorted(set(np.bincount(ch_s).tolist()))
        kb, pb = k // 8, (n - k) // 8
        for data in (rng.integers(0, 256, kb, dtype=np.uint8), np.zeros(kb, np.uint8), (np.arange(kb) % 256).astype(np.uint8)):
            pt = np.zeros(pb, np.uint8)
            assert L.labrador_ldpc_host_encode_model(code, data.ctypes.data, pt.ctypes.data, None) == 1
            tx = np.unpackbits(np.concatenate([data, pt]))                # the n transmitted bits
            # punctured bits: row 2 is [data terms] + S(CB) + I(CC), so bit j of CC = parity of the row-2 check over the rest
            bits = np.concatenate([tx, np.zeros(p, np.uint8)])
            row2 = (checks >= 2 * m) & (vars_ < n)
            par2 = np.bincount(checks[row2] - 2 * m, weights=bits[vars_[row2]], minlength=m).astype(np.int64) & 1
            bits[n:] = par2
            syn = np.bincount(checks, weights=bits[vars_], minlength=n + p - k).astype(np.int64) & 1
            assert not syn.any(), (code, int(syn.sum()))
            if not data.any():
                assert not pt.any()


def test_header_macros_compile_and_match(kats, tmp_path):
    # compile a C program against include/labrador_ldpc.h and compare every size macro
    # (incl. the reference's misspelt _TM6140 names) with the reference's literal table
    lines = ['#include "labrador_ldpc.h"', "#include <stdio.h>", "int main(void){"]
    for name in NAMES + ["TM6140"]:
        for mac in ("N", "K", "BF_WORKING_LEN", "MS_WORKING_LEN", "MS_WORKING_U8_LEN", "OUTPUT_LEN"):
            lines.append('printf("%s_%s %%d\\n", (int)LABRADOR_LDPC_%s_%s);' % (mac, name, mac, name))
    lines.append('printf("GEN %d %d\\n", (int)LABRADOR_LDPC_MS_WORKING_LEN(TM8192), (int)LABRADOR_LDPC_CODE(TM2048));')
    lines.append("return 0;}")
    src = tmp_path / "m.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "m"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines() if not l.startswith("GEN"))
    field = {"N": "n", "K": "k", "BF_WORKING_LEN": "decode_bf_working_len", "MS_WORKING_LEN": "decode_ms_working_len",
             "MS_WORKING_U8_LEN": "decode_ms_working_u8_len", "OUTPUT_LEN": "output_len"}
    for name in NAMES + ["TM6140"]:
        P = kats["params"]["TM6144" if name == "TM6140" else name]
        for mac, f in field.items():
            assert int(out["%s_%s" % (mac, name)]) == P[f], (mac, name)
    # every macro the reference header defines exists here with the same value, except the
    # reference's N_TM6144 = 6140 typo which is corrected to 6144 (documented in the header)
    for mname, val in kats["header_macros"].items():
        key = mname.replace("LABRADOR_LDPC_", "")
        if mname == "LABRADOR_LDPC_N_TM6144":
            assert val == 6140 and int(out["N_TM6144"]) == 6144
        else:
            assert int(out[key]) == val, mname


def test_example_c_builds_against_header(tmp_path):
    # the reference's capi/examples/example.c pattern: static buffers sized by the macros
    src = tmp_path / "ex.c"
    src.write_text('''
#include <stdint.h>
#include <stdbool.h>
#include "labrador_ldpc.h"
#define CODE TC128
uint8_t message[LABRADOR_LDPC_K(CODE)/8];
uint8_t codeword[LABRADOR_LDPC_N(CODE)/8];
float llrs[LABRADOR_LDPC_N(CODE)];
float working[LABRADOR_LDPC_MS_WORKING_LEN(CODE)];
uint8_t working_u8[LABRADOR_LDPC_MS_WORKING_U8_LEN(CODE)];
uint8_t output[LABRADOR_LDPC_OUTPUT_LEN(CODE)];
int main(void) {
    enum labrador_ldpc_code code = LABRADOR_LDPC_CODE(CODE);
    if (labrador_ldpc_code_n(code) != 128) return 1;
    if (sizeof(working)/sizeof(working[0]) != labrador_ldpc_ms_working_len(code)) return 2;
    if (sizeof(output) != labrador_ldpc_output_len(code)) return 3;
    return 0;
}''')
    import labrador_ldpc_b200 as L
    libdir = os.path.dirname(L._LIB_PATH)
    exe = tmp_path / "ex"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-llabrador_ldpc", "-Wl,-rpath," + libdir])
    assert subprocess.call([str(exe)]) == 0


def build_example(name, outdir):
    import labrador_ldpc_b200 as L
    libdir = os.path.dirname(L._LIB_PATH)
    exe = os.path.join(str(outdir), name)
    subprocess.check_call(["gcc", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", name + ".c"),
                           "-I", os.path.join(ROOT, "include"), "-L", libdir, "-llabrador_ldpc", "-lm",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_examples_compile_as_plain_c(tmp_path):
    """examples/example.c (the reference's capi example pattern) and examples/batch_example.c build with a
    C compiler against include/labrador_ldpc.h alone; they are run on the GPU box by tests/test_gpu_parity.py."""
    for name in ("example", "batch_example"):
        assert os.path.exists(build_example(name, tmp_path))


def test_fresh_checkout_builds_on_import(ldpc, tmp_path):
    """A checkout has no lib/ (the .so is git-ignored): importing the package -- which is what
    __graft_entry__.build() does first -- must run the in-tree build instead of failing.  Checked on a copy of
    the package whose build step is replaced by "copy the real .so" so the test stays fast."""
    import shutil
    pkg = tmp_path / "labrador_ldpc_b200"
    shutil.copytree(os.path.join(ROOT, "labrador_ldpc_b200"), pkg,
                    ignore=shutil.ignore_patterns("lib", "build", "__pycache__"))
    shutil.copytree(os.path.join(ROOT, "include"), tmp_path / "include")
    assert not (pkg / "lib").exists()
    with open(pkg / "_build.py", "a") as f:
        f.write("\n\ndef build(force=False, verbose=False):\n"
                "    os.makedirs(LIBDIR, exist_ok=True)\n"
                "    shutil.copy(%r, LIB)\n"
                "    open(LIB + '.stamp', 'w').write(_stamp())\n"
                "    return LIB\n" % ldpc._LIB_PATH)
    out = subprocess.check_output([sys.executable, "-c",
                                   "import sys; sys.path.insert(0, %r); import labrador_ldpc_b200 as L; "
                                   "print(L._LIB_PATH); print(L.version())" % str(tmp_path)], text=True)
    assert str(pkg / "lib") in out and "labrador-ldpc-b200" in out


def test_argument_validation_needs_no_gpu(ldpc):
    L = ldpc.lib
    buf = np.zeros(64, np.uint8)
    p = buf.ctypes.data
    assert L.labrador_ldpc_decode_ms_i8_batch(12, p, p, 1, 10, p, None) == -1     # bad code
    assert L.labrador_ldpc_decode_ms_i8_batch(0, None, p, 1, 10, p, None) == -2   # null pointer
    assert L.labrador_ldpc_decode_ms_i8_batch(0, None, None, 0, 10, None, None) == 0  # empty batch is a no-op
    assert L.labrador_ldpc_decode_bf_batch(-1, p, p, 1, 10, p, None) == -1
    assert L.labrador_ldpc_copy_encode_batch(0, None, p, 1) == -2
    assert L.labrador_ldpc_decode_ms_batch_async(0, 7, p, p, 1, 10, p, None, None) == -5   # bad llr type
    assert b"llr_type" in L.labrador_ldpc_last_error()
    # fused front ends: limit / scale / (front, type) combinations are checked before any pointer is touched
    assert L.labrador_ldpc_decode_ms_i8_soft_batch(0, p, 4.0, 0, p, 1, 10, p, None) == -5
    assert L.labrador_ldpc_decode_ms_i8_soft_batch(0, p, 4.0, 128, p, 1, 10, p, None) == -5
    assert L.labrador_ldpc_decode_ms_i16_soft_batch(0, p, 4.0, 32768, p, 1, 10, p, None) == -5
    assert L.labrador_ldpc_decode_ms_i8_soft_batch(0, p, float("inf"), 31, p, 1, 10, p, None) == -5
    assert L.labrador_ldpc_decode_ms_i8_soft_batch(0, p, float("nan"), 31, p, 1, 10, p, None) == -5
    assert L.labrador_ldpc_decode_ms_i8_soft_batch(12, p, 4.0, 31, p, 1, 10, p, None) == -1
    assert L.labrador_ldpc_decode_ms_i8_soft_batch(0, None, 4.0, 31, p, 1, 10, p, None) == -2
    assert L.labrador_ldpc_decode_ms_i8_soft_batch(0, None, 4.0, 31, None, 0, 10, None, None) == 0
    assert L.labrador_ldpc_decode_ms_front_batch_async(0, 3, 1, p, 4.0, 31, p, 1, 10, p, None, None) == -5   # soft -> f32
    assert L.labrador_ldpc_decode_ms_front_batch_async(0, 1, 2, p, 4.0, 31, p, 1, 10, p, None, None) == -5   # hard -> i16
    assert L.labrador_ldpc_decode_ms_front_batch_async(0, 0, 3, p, 4.0, 31, p, 1, 10, p, None, None) == -5   # bad front
    assert L.labrador_ldpc_quantise_i8_batch(0, p, 4.0, 200, p, 1) == -5
    assert L.labrador_ldpc_quantise_batch_async(0, 3, p, 4.0, 31, p, 1, None) == -5


def test_philox_restatement_known_answers():
    """Random123 known-answer vectors for philox4x32-10; pins tests/frames.py::philox4x32_10, against which the
    device generator (csrc/channel.cu) is compared bit for bit in tests/test_gpu_channel.py."""
    from frames import philox4x32_10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert tuple(int(x) for x in got) == want


def test_harness_kernel_argument_validation(ldpc):
    L = ldpc.lib
    buf = np.zeros(64, np.uint8)
    p = buf.ctypes.data
    assert L.labrador_ldpc_random_data_batch(12, 1, 0, p, 1) == -1
    assert L.labrador_ldpc_random_data_batch(0, 1, 0, None, 1) == -2
    assert L.labrador_ldpc_random_data_batch(0, 1, 0, None, 0) == 0
    assert L.labrador_ldpc_awgn_batch(0, 2, p, 1.0, 1.0, 31, 1, 0, p, 1) == -5          # i32 output
    assert L.labrador_ldpc_awgn_batch(0, 0, p, 1.0, 1.0, 0, 1, 0, p, 1) == -5           # limit
    assert L.labrador_ldpc_awgn_batch(0, 3, p, -1.0, 1.0, 0, 1, 0, p, 1) == -5          # sigma < 0
    assert L.labrador_ldpc_awgn_batch(0, 3, p, float("nan"), 1.0, 0, 1, 0, p, 1) == -5
    assert L.labrador_ldpc_awgn_batch(0, 3, None, 1.0, 1.0, 0, 1, 0, p, 1) == -2
    assert L.labrador_ldpc_count_errors_batch(0, p, p, None, 1) == -2
    assert L.labrador_ldpc_count_errors_batch(12, p, p, p, 1) == -1


def test_python_mirror_length_asserts(ldpc):
    # the reference asserts every buffer length (src/decoder.rs:356-359, src/encoder.rs:296,312-313)
    c = ldpc.LDPCCode.TC128
    with pytest.raises(ValueError):
        c.decode_ms(np.zeros(127, np.int8), np.zeros(16, np.uint8))
    with pytest.raises(ValueError):
        c.decode_ms(np.zeros(128, np.int8), np.zeros(15, np.uint8))
    with pytest.raises(ValueError):
        c.decode_ms(np.zeros(128, np.int8), np.zeros(16, np.uint8), working=np.zeros(3, np.int8))
    with pytest.raises(ValueError):
        c.decode_bf(np.zeros(15, np.uint8), np.zeros(16, np.uint8))
    with pytest.raises(ValueError):
        c.copy_encode(np.zeros(7, np.uint8), np.zeros(16, np.uint8))
    with pytest.raises(ValueError):
        c.encode(np.zeros(15, np.uint8))


def test_no_cpu_fallback_without_gpu(ldpc):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = ldpc.LDPCCode.TC128
    with pytest.raises(ldpc.LdpcError):
        c.decode_ms_batch(np.zeros((2, 128), np.int8), 10)
    with pytest.raises(ldpc.LdpcError):
        c.copy_encode_batch(np.zeros((2, 8), np.uint8))


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "labrador_ldpc_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text.replace(
                    "oracle/ccsds_tables.h", ""), (dirpath, f)


def test_rust_binding_matches_header(ldpc):
    """rust/labrador-ldpc-b200 cannot be compiled here (no Rust toolchain); at least every
    extern it declares must exist in the header and be exported by the shared library."""
    src = open(os.path.join(ROOT, "rust", "labrador-ldpc-b200", "src", "lib.rs")).read()
    externs = set(re.findall(r"pub fn (labrador_ldpc_[a-z0-9_]+)\s*\(", src))
    assert len(externs) >= 25
    declared = set(header_symbols())
    for name in externs:
        assert name in declared, name
        assert hasattr(ldpc.lib, name), name
    # enum discriminants are the reference's (src/codes/mod.rs:37-66)
    for i, name in enumerate(NAMES):
        assert re.search(r"\b%s = %d\b" % (name, i), src), name


def test_batch_result_arrays_are_size_checked(ldpc):
    """The C side writes `batch` entries into output / success / iters: the Python mirror must reject short
    caller-supplied arrays before the call (no GPU needed: the check comes first)."""
    c = ldpc.LDPCCode.TC128
    llrs = np.zeros((3, 128), np.int8)
    hard = np.zeros((3, 16), np.uint8)
    ok3, it3, out3 = np.zeros(3, np.uint8), np.zeros(3, np.uint32), np.zeros((3, 16), np.uint8)
    for kwargs in ({"success": np.zeros(2, np.uint8)}, {"iters": np.zeros(2, np.uint32)},
                   {"output": np.zeros((2, 16), np.uint8)}, {"iters": np.zeros(3, np.uint16)}):
        args = {"output": out3, "success": ok3, "iters": it3}
        args.update(kwargs)
        with pytest.raises(ValueError):
            c.decode_ms_batch(llrs, 10, **args)
        with pytest.raises(ValueError):
            c.decode_bf_batch(hard, 10, **args)
        with pytest.raises(ValueError):
            c.decode_ms_hard_batch(hard, 10, **args)
    with pytest.raises(ValueError):
        c.hard_to_llrs_batch(hard, "i8", llrs=np.zeros((2, 128), np.int8))
    with pytest.raises(ValueError):
        c.llrs_to_hard_batch(llrs, output=np.zeros((2, 16), np.uint8))


def test_rust_batch_functions_check_every_slice():
    """Every safe `_batch` wrapper that hands success / iters to C asserts their lengths first."""
    src = open(os.path.join(ROOT, "rust", "labrador-ldpc-b200", "src", "lib.rs")).read()
    bodies = re.findall(r"pub fn (\w+_batch)[^{]*\{(.*?)\n    \}", src, flags=re.S)
    assert len(bodies) >= 4
    for name, body in bodies:
        if "success.as_mut_ptr()" in body:
            assert "assert_eq!(success.len(), batch)" in body, name
            assert "assert_eq!(iters.len(), batch)" in body, name
