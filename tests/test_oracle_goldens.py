"""Pins the CPU oracle against every golden the reference's own tests hold.

Mirrors: src/codes/mod.rs:517-535, src/encoder.rs:361-527, src/lib.rs:21-50,130-144,
src/decoder.rs:531-699.
"""
import numpy as np
import pytest

CODES = list(range(9))
NAMES = ["TC128", "TC256", "TC512", "TM1280", "TM1536", "TM2048", "TM5120", "TM6144", "TM8192"]


def txdata(oracle, code):
    return (np.arange(oracle.k(code) // 8) % 256).astype(np.uint8)


@pytest.mark.parametrize("code", CODES)
def test_iter_parity_crc(oracle, kats, code):
    # src/codes/mod.rs:517-535
    cnt, checks, vars_, crc = oracle.edges(code)
    assert cnt == kats["params"][NAMES[code]]["paritycheck_sum"]
    assert crc == kats["edge_crc32"][code]
    n, k, p = oracle.n(code), oracle.k(code), oracle.p(code)
    assert checks.max() == n + p - k - 1
    assert vars_.max() == n + p - 1


@pytest.mark.parametrize("code", CODES)
def test_length_formulas(oracle, kats, code):
    # src/decoder.rs:531-551 and capi/include/labrador_ldpc.h:45-113
    P = kats["params"][NAMES[code]]
    assert oracle.n(code) == P["n"] and oracle.k(code) == P["k"]
    assert oracle.p(code) == P["punctured_bits"]
    assert oracle.m(code) == P["submatrix_size"] and oracle.b(code) == P["circulant_size"]
    assert oracle.bf_working_len(code) == P["decode_bf_working_len"]
    assert oracle.ms_working_len(code) == P["decode_ms_working_len"]
    assert oracle.ms_working_u8_len(code) == P["decode_ms_working_u8_len"]
    assert oracle.output_len(code) == P["output_len"]
    M = kats["header_macros"]
    name = NAMES[code]
    assert M["LABRADOR_LDPC_K_" + name] == P["k"]
    if name != "TM6144":   # the reference header's TM6144 typos (6140) are documented in SURVEY.md
        assert M["LABRADOR_LDPC_N_" + name] == P["n"]
        assert M["LABRADOR_LDPC_MS_WORKING_LEN_" + name] == P["decode_ms_working_len"]
        assert M["LABRADOR_LDPC_OUTPUT_LEN_" + name] == P["output_len"]


@pytest.mark.parametrize("code", CODES)
def test_encode_kat(oracle, kats, code):
    # src/encoder.rs:324-359: u8, u32 and u64 paths all give the known parity block
    data = txdata(oracle, code)
    want = np.array(kats["encode_parity"][NAMES[code]], np.uint8)
    for word in (8, 32, 64):
        cw = oracle.copy_encode(code, data, word)
        assert np.array_equal(cw[: data.size], data)
        assert np.array_equal(cw[data.size:], want), (NAMES[code], word)


def test_doctest_encode_views(oracle, kats):
    # src/lib.rs:130-144
    cw = oracle.copy_encode(0, np.arange(8, dtype=np.uint8), 32)
    assert cw.tolist() == kats["doctest_tc128_codeword"]
    cw64 = oracle.copy_encode(0, np.arange(8, dtype=np.uint8), 64)
    assert int(cw64[8:].view("<u8")[0]) == kats["doctest_tc128_u64_parity_word"]


def test_doctest_decode_bf(oracle):
    # src/lib.rs:21-50
    data = np.arange(8, dtype=np.uint8)
    cw = oracle.copy_encode(0, data)
    rx = cw.copy()
    rx[0] ^= 0x55
    ok, iters, out = oracle.decode_bf(0, rx, 20)
    assert np.array_equal(out[:8], data)


def test_converters(oracle, kats):
    # src/decoder.rs:553-605
    hard = np.array(kats["convert_hard"], np.uint8)
    want = np.array(kats["convert_llrs"], np.float32)
    for ty in ("i8", "i16", "i32", "f32", "f64"):
        llrs = oracle.hard_to_llrs(0, hard, ty)
        assert np.array_equal(llrs.astype(np.float32), want)
        assert np.array_equal(oracle.llrs_to_hard(0, llrs, ty), hard)
    # -0.0 is not a 1 bit (hard_bit is `< 0`, src/decoder.rs:76)
    z = np.full(128, -0.0, np.float32)
    assert not oracle.llrs_to_hard(0, z).any()


@pytest.mark.parametrize("code", [c for c in CODES if c >= 3])
def test_decode_erasures(oracle, code):
    # src/decoder.rs:607-645
    cw = oracle.copy_encode(code, txdata(oracle, code))
    full = np.zeros(oracle.output_len(code), np.uint8)
    full[: cw.size] = cw
    ok, iters, out = oracle.decode_erasures(code, full, 50)
    assert ok
    llrs = oracle.hard_to_llrs(code, cw, "i8")
    ok2, _, out_ms = oracle.decode_ms(code, llrs, 50)
    assert ok2
    assert np.array_equal(out, out_ms)
    assert iters == 0     # SURVEY.md section 7 hard part 7: always exactly one pass


@pytest.mark.parametrize("code", CODES)
def test_decode_bf(oracle, code):
    # src/decoder.rs:647-670
    cw = oracle.copy_encode(code, txdata(oracle, code))
    rx = cw.copy()
    rx[0] ^= (1 << 7) | (1 << 5) | (1 << 3)
    ok, iters, out = oracle.decode_bf(code, rx, 50)
    assert ok
    assert np.array_equal(out[: cw.size], cw)
    assert iters == (1 if code < 3 else 2)   # derived sanity values, SURVEY.md section 4


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("ty", ["i8", "i16", "i32", "f32", "f64"])
def test_decode_ms(oracle, code, ty):
    # src/decoder.rs:671-699 (the reference only runs i8; the other types use the same vector)
    cw = oracle.copy_encode(code, txdata(oracle, code))
    rx = cw.copy()
    rx[0] ^= (1 << 7) | (1 << 5) | (1 << 3)
    llrs = oracle.hard_to_llrs(code, rx, ty)
    ok, iters, out = oracle.decode_ms(code, llrs, 50)
    assert ok
    assert np.array_equal(out[: cw.size], cw)
    assert iters == (2 if code < 3 else 3)   # derived sanity values, SURVEY.md section 4


@pytest.mark.parametrize("code", CODES)
def test_decode_ms_maxiters_zero_and_failure(oracle, code):
    # SURVEY.md section 7 hard part 6: maxiters == 0 -> (false, 0) and all-zero output
    llrs = np.full(oracle.n(code), -3, np.int8)
    ok, iters, out = oracle.decode_ms(code, llrs, 0)
    assert (ok, iters) == (False, 0) and not out.any()
    # garbage LLRs: must report failure with iters == maxiters
    rng = np.random.default_rng(5)
    llrs = rng.integers(-128, 128, oracle.n(code)).astype(np.int8)
    ok, iters, out = oracle.decode_ms(code, llrs, 7)
    assert (ok, iters) == (False, 7)


def test_batch_matches_single(oracle):
    rng = np.random.default_rng(1)
    code = 5
    cw = oracle.copy_encode_batch(code, rng.integers(0, 256, (6, 128)).astype(np.uint8), nthreads=2)
    bits = np.unpackbits(cw, axis=1).astype(np.float32)
    y = (1 - 2 * bits) + rng.normal(0, 0.7, bits.shape).astype(np.float32)
    llrs = np.clip(np.rint(8 * y), -31, 31).astype(np.int8)
    out, succ, iters = oracle.decode_ms_batch(code, llrs, 30, nthreads=3)
    for f in range(6):
        ok, it, o = oracle.decode_ms(code, llrs[f], 30)
        assert ok == bool(succ[f]) and it == iters[f] and np.array_equal(o, out[f])
