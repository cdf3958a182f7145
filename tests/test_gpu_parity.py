"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Integer LLR types must match the oracle exactly (decoded bytes, success flag, iteration
count) on every frame; float types must give identical hard decisions on >= 99.99 % of
frames (north_star) -- tolerance written in `assert_float_parity`.
"""
import ctypes
import os

import numpy as np
import pytest

from frames import hard_frames, make_frames

pytestmark = pytest.mark.gpu

NAMES = ["TC128", "TC256", "TC512", "TM1280", "TM1536", "TM2048", "TM5120", "TM6144", "TM8192"]
CODES = list(range(9))
# Eb/N0 per code where decoding mostly, but not always, succeeds (mix of iteration counts)
EBN0 = {0: 3.0, 1: 3.0, 2: 2.5, 3: 3.6, 4: 2.6, 5: 1.8, 6: 3.4, 7: 2.4, 8: 1.6}


def assert_exact(got, want, what):
    out_g, ok_g, it_g = got
    out_w, ok_w, it_w = want
    assert np.array_equal(np.asarray(ok_g).astype(bool), np.asarray(ok_w).astype(bool)), what + ": success flags differ"
    assert np.array_equal(np.asarray(it_g).astype(np.int64), np.asarray(it_w).astype(np.int64)), what + ": iteration counts differ"
    assert np.array_equal(np.asarray(out_g), np.asarray(out_w)), what + ": decoded bytes differ"


def assert_float_parity(got, want, what, min_frac=0.9999):
    out_g, ok_g, it_g = got
    out_w, ok_w, it_w = want
    same = (np.asarray(out_g) == np.asarray(out_w)).all(axis=1)
    same &= np.asarray(ok_g).astype(bool) == np.asarray(ok_w).astype(bool)
    same &= np.asarray(it_g).astype(np.int64) == np.asarray(it_w).astype(np.int64)
    assert same.mean() >= min_frac, "%s: only %.5f of frames identical" % (what, same.mean())


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("ty", ["i8", "i16", "i32", "f32", "f64"])
def test_reference_decode_ms_vector(ldpc, oracle, code, ty):
    # reference src/decoder.rs:671-699 for every LLR type, single-codeword C entry point semantics
    c = ldpc.LDPCCode(code)
    data = (np.arange(c.k() // 8) % 256).astype(np.uint8)
    cw = np.zeros(c.n() // 8, np.uint8)
    c.copy_encode(data, cw)
    assert np.array_equal(cw, oracle.copy_encode(code, data))
    rx = cw.copy()
    rx[0] ^= (1 << 7) | (1 << 5) | (1 << 3)
    llrs = np.zeros(c.n(), ldpc._NP_OF[ty])
    c.hard_to_llrs(rx, llrs)
    assert np.array_equal(llrs, oracle.hard_to_llrs(code, rx, ty))
    out = np.zeros(c.output_len(), np.uint8)
    ok, iters = c.decode_ms(llrs, out, maxiters=50)
    ok_w, it_w, out_w = oracle.decode_ms(code, llrs, 50)
    assert ok and ok_w and iters == it_w
    assert np.array_equal(out, out_w)
    assert np.array_equal(out[: cw.size], cw)


@pytest.mark.parametrize("code", CODES)
def test_reference_decode_bf_vector(ldpc, oracle, code):
    # reference src/decoder.rs:647-670 and the doc-test src/lib.rs:21-50
    c = ldpc.LDPCCode(code)
    data = (np.arange(c.k() // 8) % 256).astype(np.uint8)
    cw = oracle.copy_encode(code, data)
    for flip in (0xA8, 0x55):
        rx = cw.copy()
        rx[0] ^= flip
        out = np.zeros(c.output_len(), np.uint8)
        ok, iters = c.decode_bf(rx, out, maxiters=50)
        ok_w, it_w, out_w = oracle.decode_bf(code, rx, 50)
        assert (ok, iters) == (ok_w, it_w)
        assert np.array_equal(out, out_w)
        if flip == 0xA8:
            assert ok and np.array_equal(out[: cw.size], cw)


def test_reference_signature_c_functions(ldpc, oracle, kats):
    """The 21 reference entry points called exactly as capi/examples/example.c does."""
    L = ldpc.lib
    code = 0
    msg = np.arange(8, dtype=np.uint8)
    cw = np.zeros(16, np.uint8)
    L.labrador_ldpc_copy_encode(code, msg.ctypes.data, cw.ctypes.data)
    assert cw.tolist() == kats["doctest_tc128_codeword"]
    cw2 = np.zeros(16, np.uint8)
    cw2[:8] = msg
    L.labrador_ldpc_encode(code, cw2.ctypes.data)
    assert np.array_equal(cw, cw2)
    # unaligned codeword pointer (the reference takes its u8 path, capi/src/lib.rs:27-33)
    raw = np.zeros(17, np.uint8)
    raw[1:9] = msg
    L.labrador_ldpc_encode(code, raw.ctypes.data + 1)
    assert np.array_equal(raw[1:], cw)
    rx = cw.copy()
    rx[7] = 0          # example.c erases the last byte of user data
    llrs = np.zeros(128, np.float32)
    L.labrador_ldpc_hard_to_llrs_f32(code, rx.ctypes.data, llrs.ctypes.data)
    assert np.array_equal(llrs, oracle.hard_to_llrs(code, rx, "f32"))
    out = np.zeros(16, np.uint8)
    working = np.zeros(L.labrador_ldpc_ms_working_len(code), np.float32)
    working_u8 = np.zeros(L.labrador_ldpc_ms_working_u8_len(code), np.uint8)
    iters = ctypes.c_size_t(99)
    ok = L.labrador_ldpc_decode_ms_f32(code, llrs.ctypes.data, out.ctypes.data, working.ctypes.data,
                                       working_u8.ctypes.data, 200, ctypes.addressof(iters))
    ok_w, it_w, out_w = oracle.decode_ms(code, llrs, 200)
    assert bool(ok) == ok_w and iters.value == it_w and np.array_equal(out, out_w)
    # iters_run may be NULL (capi/src/lib.rs:77,91)
    ok = L.labrador_ldpc_decode_ms_f32(code, llrs.ctypes.data, out.ctypes.data, working.ctypes.data,
                                       working_u8.ctypes.data, 200, None)
    assert bool(ok) == ok_w
    hard = np.zeros(16, np.uint8)
    L.labrador_ldpc_llrs_to_hard_f32(code, llrs.ctypes.data, hard.ctypes.data)
    assert np.array_equal(hard, rx)
    bfw = np.zeros(L.labrador_ldpc_bf_working_len(code), np.uint8)
    rx2 = cw.copy()
    rx2[0] ^= 0x55
    ok = L.labrador_ldpc_decode_bf(code, rx2.ctypes.data, out.ctypes.data, bfw.ctypes.data, 20, ctypes.addressof(iters))
    ok_w, it_w, out_w = oracle.decode_bf(code, rx2, 20)
    assert bool(ok) == ok_w and iters.value == it_w and np.array_equal(out, out_w)
    for t, npdt in (("i8", np.int8), ("i16", np.int16), ("f64", np.float64)):
        l2 = np.zeros(128, npdt)
        getattr(L, "labrador_ldpc_hard_to_llrs_" + t)(code, rx.ctypes.data, l2.ctypes.data)
        assert np.array_equal(l2, oracle.hard_to_llrs(code, rx, t))
        h2 = np.zeros(16, np.uint8)
        getattr(L, "labrador_ldpc_llrs_to_hard_" + t)(code, l2.ctypes.data, h2.ctypes.data)
        assert np.array_equal(h2, rx)


def test_converter_kat(ldpc, kats):
    # reference src/decoder.rs:553-605
    c = ldpc.LDPCCode.TC128
    hard = np.array(kats["convert_hard"], np.uint8)
    want = np.array(kats["convert_llrs"], np.float32)
    llrs = np.zeros(128, np.float32)
    c.hard_to_llrs(hard, llrs)
    assert np.array_equal(llrs, want)
    back = np.zeros(16, np.uint8)
    c.llrs_to_hard(llrs, back)
    assert np.array_equal(back, hard)
    z = np.full(128, -0.0, np.float32)     # -0.0 is a 0 bit
    c.llrs_to_hard(z, back)
    assert not back.any()


@pytest.mark.parametrize("code", CODES)
def test_encode_kat_and_random(ldpc, oracle, kats, code):
    # reference src/encoder.rs:361-527 known answers, then random data against the oracle
    c = ldpc.LDPCCode(code)
    data = (np.arange(c.k() // 8) % 256).astype(np.uint8)
    cw = np.zeros(c.n() // 8, np.uint8)
    c.copy_encode(data, cw)
    assert cw[data.size:].tolist() == kats["encode_parity"][NAMES[code]]
    rng = np.random.default_rng(code)
    batch = 257
    d = rng.integers(0, 256, (batch, c.k() // 8), dtype=np.uint8)
    d[0] = 0
    d[1] = 0xFF
    want = oracle.copy_encode_batch(code, d, nthreads=4)
    got = c.copy_encode_batch(d)
    assert np.array_equal(got, want)
    inplace = np.zeros_like(want)
    inplace[:, : c.k() // 8] = d
    c.encode_batch(inplace)
    assert np.array_equal(inplace, want)
    # linearity over GF(2): enc(a ^ b) == enc(a) ^ enc(b)
    x = c.copy_encode_batch(d[2:3] ^ d[3:4])
    assert np.array_equal(x[0], want[2] ^ want[3])
    # ragged batches: the TM encoder packs up to eight codewords into a warp
    for b in (1, 2, 3, 5, 8, 9, 31, 33):
        assert np.array_equal(c.copy_encode_batch(d[:b]), want[:b]), b


@pytest.mark.parametrize("code", CODES)
def test_encode_large_batch_and_unaligned(ldpc, oracle, code):
    """Large batches take the lookup-table encoders (TM: encode_tm.cu, TC: encode_tc_lut_kernel), small ones the
    compact TM form / the generator kernel; also byte-granular pointers (plain-load paths) and the in-place
    form.  All against the oracle."""
    import torch
    c = ldpc.LDPCCode(code)
    kb, nb = c.k() // 8, c.n() // 8
    batch = 200001 if code < 3 else (40000 if code in (3, 4, 5) else 12001)
    rng = np.random.default_rng(100 + code)
    d = rng.integers(0, 256, (batch, kb), dtype=np.uint8)
    want = oracle.copy_encode_batch(code, d, nthreads=os.cpu_count() or 1)
    dev = torch.from_numpy(d).cuda()
    got = c.copy_encode_batch(dev)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), want)
    inplace = torch.zeros((batch, nb), dtype=torch.uint8, device="cuda")
    inplace[:, :kb] = dev
    c.encode_batch(inplace)
    torch.cuda.synchronize()
    assert np.array_equal(inplace.cpu().numpy(), want)
    # unaligned device buffers (offset 1 and 2 bytes); the TC table kernel only runs on large batches
    small = batch if code < 3 else 67
    src = torch.zeros(small * kb + 8, dtype=torch.uint8, device="cuda")
    dst = torch.zeros(small * nb + 8, dtype=torch.uint8, device="cuda")
    src[1:1 + small * kb] = dev[:small].reshape(-1)
    c.copy_encode_batch(src[1:1 + small * kb].view(small, kb), dst[2:2 + small * nb].view(small, nb))
    torch.cuda.synchronize()
    assert np.array_equal(dst[2:2 + small * nb].view(small, nb).cpu().numpy(), want[:small])
    assert int(dst[:2].sum()) == 0 and int(dst[2 + small * nb:].sum()) == 0


def test_small_call_staging_boundary(ldpc, oracle):
    """Host-pointer calls of up to 256 KB go through the pinned, device-mapped staging block, larger ones through
    the chunk pipeline (csrc/runtime.cu); batches on both sides of the boundary must agree with the oracle."""
    code = 5
    c = ldpc.LDPCCode(code)
    _, _, llrs = make_frames(oracle, code, 130, 2.4, seed=321, ty="i8")
    want = oracle.decode_ms_batch(code, llrs, 60, nthreads=8)
    for batch in (1, 2, 105, 108, 109, 110, 111, 112, 130):     # 256 KB / (2048 + 320 + 1 + 4 bytes, 16-byte rounded) = 110.x frames
        got = c.decode_ms_batch(llrs[:batch], 60)
        assert_exact(got, tuple(w[:batch] for w in want), "batch %d" % batch)
    data = np.random.default_rng(5).integers(0, 256, (700, c.k() // 8), dtype=np.uint8)
    cw = oracle.copy_encode_batch(code, data, nthreads=4)
    for batch in (1, 680, 682, 683, 684, 700):                   # 256 KB / (128 + 256) = 682.6 frames
        assert np.array_equal(c.copy_encode_batch(data[:batch]), cw[:batch]), batch
        ip = np.zeros((batch, c.n() // 8), np.uint8)
        ip[:, : c.k() // 8] = data[:batch]
        c.encode_batch(ip)
        assert np.array_equal(ip, cw[:batch]), batch


def test_concurrent_host_threads(ldpc, oracle):
    """The reference is re-entrant and its perftest calls it from every core at once
    (perftest/src/main.rs:39-45); the drop-in must give every thread its own correct answer."""
    import threading
    jobs = []
    for t in range(8):
        code = (0, 2, 5, 8, 3, 1, 6, 7)[t]
        data, cw, llrs = make_frames(oracle, code, 6, EBN0[code] + 1.0, seed=700 + t, ty="i8")
        jobs.append((code, data, cw, llrs, oracle.decode_ms_batch(code, llrs, 50, nthreads=2)))
    errors = []

    def worker(job):
        code, data, cw, llrs, want = job
        c = ldpc.LDPCCode(code)
        try:
            for rep in range(5):
                for f in range(data.shape[0]):                   # single-codeword reference-signature calls
                    out = np.zeros(c.n() // 8, np.uint8)
                    c.copy_encode(data[f], out)
                    assert np.array_equal(out, cw[f])
                    dec = np.zeros(c.output_len(), np.uint8)
                    ok, it = c.decode_ms(llrs[f], dec, maxiters=50)
                    assert np.array_equal(dec, want[0][f]) and bool(ok) == bool(want[1][f]) and int(it) == int(want[2][f])
                got = c.decode_ms_batch(llrs, 50)                # and a batched call from host memory
                assert_exact(got, want, NAMES[code] + " threaded")
                assert np.array_equal(c.copy_encode_batch(data), cw)
        except Exception as e:                                   # noqa: BLE001
            errors.append((code, repr(e)))

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("knob", ["LABRADOR_LDPC_ENC_GENERATOR=1", "LABRADOR_LDPC_ENC_TM_FORM=1", "LABRADOR_LDPC_ENC_TC_COPIES=0", "LABRADOR_LDPC_ENC_TC_COPIES=1"])
def test_alternative_encoder_kernels_stay_exact(knob):
    """The default encoders are the table kernels (TM: through the parity-check matrix with a nibble table,
    encode_tm.cu; TC: per-nibble / per-byte tables).  The generator kernel, the compact TM form and the TC table
    kernels with / without bank-spreading table copies are reachable through environment knobs and must stay
    bit-exact as well."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle")
import labrador_ldpc_b200 as L, pyoracle
o = pyoracle.Oracle()
for code in range(9):
    c = L.LDPCCode(code)
    d = np.random.default_rng(code).integers(0, 256, (301, c.k() // 8), dtype=np.uint8)
    want = o.copy_encode_batch(code, d, nthreads=4)
    assert np.array_equal(c.copy_encode_batch(d), want), code
    assert np.array_equal(c.copy_encode_batch(d[:1]), want[:1]), code
    ip = np.zeros_like(want); ip[:, : c.k() // 8] = d
    c.encode_batch(ip)
    assert np.array_equal(ip, want), code
print("OK")
''' % (root, root)
    name, value = knob.split("=")
    env = dict(os.environ, **{name: value})
    out = subprocess.check_output([sys.executable, "-c", script], env=env, text=True)
    assert "OK" in out


@pytest.mark.parametrize("code", CODES)
def test_decode_ms_i8_awgn_exact(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    batch = 384 if code < 6 else 160
    _, _, llrs = make_frames(oracle, code, batch, EBN0[code], seed=100 + code, ty="i8")
    want = oracle.decode_ms_batch(code, llrs, 100, nthreads=8)
    got = c.decode_ms_batch(llrs, 100)
    assert_exact(got, want, NAMES[code] + " i8")
    assert 0 < want[1].sum(), "test vector should contain successes"


@pytest.mark.parametrize("code", [0, 1, 2])
def test_tc_packed_pair_kernel_streams(ldpc, oracle, code):
    """decode_ms_tc_x2.cu: two codewords share every register and each 16-bit half runs its own stream of
    frames.  Odd batches (one half runs dry first), a mix of frames that converge at once, slowly and never
    (re-initialisation of one half while its partner is mid-decode), tight iteration caps, and the A/B kernel."""
    c = ldpc.LDPCCode(code)
    assert c.decode_ms_kernel_name("i8") == "ms_tc_x2<i8>"
    rng = np.random.default_rng(900 + code)
    parts = [make_frames(oracle, code, 301, eb, seed=910 + 7 * code + i, ty="i8")[2] for i, eb in enumerate((0.0, 2.5, 6.0))]
    llrs = np.concatenate(parts)
    llrs = llrs[rng.permutation(llrs.shape[0])]                  # 903 frames: hopeless, marginal and clean ones interleaved
    for batch, iters in ((903, 100), (2, 100), (3, 7), (257, 1), (64, 2), (5, 100)):
        want = oracle.decode_ms_batch(code, llrs[:batch], iters, nthreads=8)
        got = c.decode_ms_batch(llrs[:batch], iters)
        assert_exact(got, want, "%s batch %d iters %d" % (NAMES[code], batch, iters))
    want = oracle.decode_ms_batch(code, llrs, 100, nthreads=8)
    assert 0 < want[1].sum() < llrs.shape[0], "needs both successes and failures"
    assert len(set(want[2].tolist())) > 5, "needs a spread of iteration counts"


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("ty", ["i16", "i32"])
def test_decode_ms_wide_int_awgn_exact(ldpc, oracle, code, ty):
    c = ldpc.LDPCCode(code)
    batch = 128 if code < 6 else 64
    _, _, llrs = make_frames(oracle, code, batch, EBN0[code], seed=200 + code, ty=ty)
    want = oracle.decode_ms_batch(code, llrs, 60, nthreads=8)
    got = c.decode_ms_batch(llrs, 60)
    assert_exact(got, want, "%s %s" % (NAMES[code], ty))


@pytest.mark.parametrize("code", [0, 1, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("ty", ["f32", "f64"])
def test_decode_ms_float_awgn(ldpc, oracle, code, ty):
    c = ldpc.LDPCCode(code)
    batch = 128 if code < 6 else 48
    _, _, llrs = make_frames(oracle, code, batch, EBN0[code], seed=300 + code, ty=ty)
    want = oracle.decode_ms_batch(code, llrs, 60, nthreads=8)
    got = c.decode_ms_batch(llrs, 60)
    assert_float_parity(got, want, "%s %s" % (NAMES[code], ty))


@pytest.mark.parametrize("code", [2, 3, 5, 6, 7, 8])
@pytest.mark.parametrize("ty", ["f32", "f64"])
def test_decode_ms_float_small_integer_llrs_exact(ldpc, oracle, code, ty):
    """Float LLRs that are small integers (with many exact zeros and some -0.0): every sum is exact, so the result
    does not depend on the order of additions and must equal the oracle's bit for bit -- while ties, zero messages
    (the `v_old == 0` arm of the self-correction rule, src/decoder.rs:422-426) and negative zeros are everywhere."""
    c = ldpc.LDPCCode(code)
    rng = np.random.default_rng(4000 + code)
    batch = 40 if code < 6 else 24
    _, cw, _ = make_frames(oracle, code, batch, 3.0, seed=50 + code, ty="i8")
    bits = np.unpackbits(cw, axis=1)[:, : c.n()].astype(np.int64)
    llrs = (1 - 2 * bits) * rng.integers(0, 4, bits.shape) + rng.integers(-1, 2, bits.shape)
    llrs = llrs.astype(np.float32 if ty == "f32" else np.float64)
    llrs[rng.random(llrs.shape) < 0.05] = -0.0
    want = oracle.decode_ms_batch(code, llrs, 25, nthreads=8)
    got = c.decode_ms_batch(llrs, 25)
    assert_exact(got, want, "%s %s" % (NAMES[code], ty))


@pytest.mark.parametrize("code", CODES)
def test_decode_ms_i8_saturation_stress(ldpc, oracle, code):
    """Full-range random LLRs: decoding fails, every saturating add/sub/abs corner is hit
    (incl. -128), and the output is the hard decision of the LAST marginals -- so the
    whole message state must match the oracle after `maxiters` iterations."""
    c = ldpc.LDPCCode(code)
    rng = np.random.default_rng(400 + code)
    batch = 64 if code < 6 else 24
    llrs = rng.integers(-128, 128, (batch, c.n())).astype(np.int8)
    llrs[0, :] = -128
    llrs[1, :] = 127
    llrs[2, ::2] = -128
    for maxiters in (1, 2, 9):
        want = oracle.decode_ms_batch(code, llrs, maxiters, nthreads=8)
        got = c.decode_ms_batch(llrs, maxiters)
        assert_exact(got, want, "%s i8 stress maxiters=%d" % (NAMES[code], maxiters))
    # large-magnitude but decodable: saturation on the way to success
    _, _, soft = make_frames(oracle, code, batch, EBN0[code] + 1.0, seed=450 + code, ty="f64")
    big = np.clip(np.rint(24.0 * soft), -127, 127).astype(np.int8)
    want = oracle.decode_ms_batch(code, big, 40, nthreads=8)
    got = c.decode_ms_batch(big, 40)
    assert_exact(got, want, NAMES[code] + " i8 big-llr")


@pytest.mark.parametrize("code", [5, 8])
def test_decode_ms_i16_saturation_stress(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    rng = np.random.default_rng(500 + code)
    llrs = rng.integers(-32768, 32768, (16, c.n())).astype(np.int16)
    want = oracle.decode_ms_batch(code, llrs, 6, nthreads=8)
    got = c.decode_ms_batch(llrs, 6)
    assert_exact(got, want, NAMES[code] + " i16 stress")


@pytest.mark.parametrize("code", [1, 3, 5])
def test_decode_ms_i32_saturation_stress(ldpc, oracle, code):
    """Full-range i32 LLRs (incl. INT32_MIN): saturating add / sub / abs corners of the 32-bit path."""
    c = ldpc.LDPCCode(code)
    rng = np.random.default_rng(550 + code)
    llrs = rng.integers(-2 ** 31, 2 ** 31, (16, c.n()), dtype=np.int64).astype(np.int32)
    llrs[0, :] = -2 ** 31
    llrs[1, :] = 2 ** 31 - 1
    llrs[2, ::2] = -2 ** 31
    for maxiters in (1, 3, 7):
        want = oracle.decode_ms_batch(code, llrs, maxiters, nthreads=8)
        got = c.decode_ms_batch(llrs, maxiters)
        assert_exact(got, want, "%s i32 stress maxiters=%d" % (NAMES[code], maxiters))


def test_decode_ms_maxiters_edge_cases(ldpc, oracle):
    for code in (0, 5):
        c = ldpc.LDPCCode(code)
        _, _, llrs = make_frames(oracle, code, 8, 4.0, seed=7, ty="i8")
        for maxiters in (0, 1):
            want = oracle.decode_ms_batch(code, llrs, maxiters)
            got = c.decode_ms_batch(llrs, maxiters)
            assert_exact(got, want, "maxiters=%d" % maxiters)
        out, ok, it = c.decode_ms_batch(llrs, 0)
        assert not out.any() and not ok.any() and not it.any()
        # clean codeword, +-1 LLRs: (true, 0) for TC, (true, 1) for TM
        cw = oracle.copy_encode(code, (np.arange(c.k() // 8) % 256).astype(np.uint8))
        l1 = oracle.hard_to_llrs(code, cw, "i8")
        out = np.zeros(c.output_len(), np.uint8)
        assert c.decode_ms(l1, out, maxiters=50) == oracle.decode_ms(code, l1, 50)[:2]


@pytest.mark.parametrize("code", CODES)
def test_decode_bf_random_errors(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    for flips in (0, 1, 3, 7, 20):
        _, _, rx = hard_frames(oracle, code, 48, flips, seed=600 + code + flips)
        want = oracle.decode_bf_batch(code, rx, 30, nthreads=8)
        got = c.decode_bf_batch(rx, 30)
        assert_exact(got, want, "%s bf flips=%d" % (NAMES[code], flips))
    for maxiters in (0, 1):
        want = oracle.decode_bf_batch(code, rx, maxiters)
        got = c.decode_bf_batch(rx, maxiters)
        assert_exact(got, want, "%s bf maxiters=%d" % (NAMES[code], maxiters))


def test_decode_bf_ragged_large_and_unaligned_batches(ldpc, oracle):
    """The TM bit-flipping kernel packs 32 / (M/32) codewords per warp and claims several groups per
    atomic on large batches: cover batches that are not a multiple of the group size, a batch large
    enough for multi-group claims, and device buffers that are not 4-byte aligned."""
    import torch
    for code in (3, 4, 5, 8):
        c = ldpc.LDPCCode(code)
        for batch in (1, 7, 45):
            _, _, rx = hard_frames(oracle, code, batch, 4, seed=650 + code + batch)
            assert_exact(c.decode_bf_batch(rx, 30), oracle.decode_bf_batch(code, rx, 30, nthreads=8),
                         "%s bf batch=%d" % (NAMES[code], batch))
        _, _, rx = hard_frames(oracle, code, 33, 5, seed=690 + code)
        want = oracle.decode_bf_batch(code, rx, 30, nthreads=8)
        buf = torch.zeros(rx.size + 8, dtype=torch.uint8, device="cuda")
        obuf = torch.zeros(33 * c.output_len() + 8, dtype=torch.uint8, device="cuda")
        for off in (1, 2):
            view = buf[off: off + rx.size].view(rx.shape)
            view.copy_(torch.from_numpy(rx))
            oview = obuf[off: off + 33 * c.output_len()].view(33, c.output_len())
            got = c.decode_bf_batch(view, 30, output=oview)
            torch.cuda.synchronize()
            assert_exact([g.cpu().numpy() for g in got], want, "%s bf unaligned off=%d" % (NAMES[code], off))
    # TC codes: batches of 2^18 frames and more are decoded in two passes (undecided frames after 6 iterations are
    # listed and redone with the full budget); frames with many errors make sure the list is exercised
    for code in (0, 1, 2):
        rng = np.random.default_rng(800 + code)
        parts = [hard_frames(oracle, code, 3000, flips, seed=810 + code + flips)[2] for flips in (0, 2, 5, 9, 14)]
        rx = np.concatenate(parts)[rng.integers(0, 15000, 300_000)]          # 300 000 frames >= the two-pass threshold
        want = oracle.decode_bf_batch(code, rx, 40, nthreads=8)
        assert (want[1] == 0).sum() >= 5 and (want[2][want[1].astype(bool)] > 6).any()   # both kinds reach the list
        assert_exact(ldpc.LDPCCode(code).decode_bf_batch(rx, 40), want, "%s bf two-pass" % NAMES[code])
        assert_exact(ldpc.LDPCCode(code).decode_bf_batch(rx, 5), oracle.decode_bf_batch(code, rx, 5, nthreads=8),
                     "%s bf maxiters below the first-pass cap" % NAMES[code])
    code, batch = 3, 100_000   # > 2 * 148 SMs * 8 warps * 4 * 8 codewords: two groups per claim
    rng = np.random.default_rng(77)
    _, _, small = hard_frames(oracle, code, 4096, 3, seed=700)
    rx = small[rng.integers(0, 4096, batch)]
    rx[:, 5] ^= rng.integers(0, 256, batch, dtype=np.uint8)       # make the frames differ
    want = oracle.decode_bf_batch(code, rx, 20, nthreads=8)
    assert_exact(ldpc.LDPCCode(code).decode_bf_batch(rx, 20), want, "TM1280 bf large batch")


@pytest.mark.parametrize("code", [c for c in CODES if c >= 3])
def test_erasure_prepass_matches_min_sum(ldpc, oracle, code):
    # reference src/decoder.rs:607-645: on a clean codeword the punctured bits recovered by
    # the erasure pre-pass of decode_bf equal those recovered by decode_ms
    c = ldpc.LDPCCode(code)
    cw = oracle.copy_encode(code, (np.arange(c.k() // 8) % 256).astype(np.uint8))
    out_bf = np.zeros(c.output_len(), np.uint8)
    ok, iters = c.decode_bf(cw, out_bf, maxiters=50)
    out_ms = np.zeros(c.output_len(), np.uint8)
    ok2, _ = c.decode_ms(oracle.hard_to_llrs(code, cw, "i8"), out_ms, maxiters=50)
    assert ok and ok2 and iters == 0
    assert np.array_equal(out_bf, out_ms)


@pytest.mark.parametrize("ty", ["i8", "i16", "i32", "f32", "f64"])
def test_converters_random(ldpc, oracle, ty):
    rng = np.random.default_rng(9)
    for code in (0, 4, 8):
        c = ldpc.LDPCCode(code)
        hard = rng.integers(0, 256, (37, c.n() // 8), dtype=np.uint8)
        llrs = c.hard_to_llrs_batch(hard, ty)
        for f in (0, 36):
            assert np.array_equal(llrs[f], oracle.hard_to_llrs(code, hard[f], ty))
        assert np.array_equal(c.llrs_to_hard_batch(llrs), hard)
        noisy = (rng.standard_normal((37, c.n())) * 50).astype(ldpc._NP_OF[ty])
        got = c.llrs_to_hard_batch(noisy)
        for f in (0, 18, 36):
            assert np.array_equal(got[f], oracle.llrs_to_hard(code, noisy[f], ty))


def test_device_pointers_pinned_and_chunking(ldpc, oracle):
    """Same frames through: pageable host, pinned host, device tensors, and a host batch
    forced through many pipeline chunks -- all must agree with the oracle."""
    import torch
    code = 5
    c = ldpc.LDPCCode(code)
    _, _, llrs = make_frames(oracle, code, 700, EBN0[code], seed=11, ty="i8")
    want = oracle.decode_ms_batch(code, llrs, 50, nthreads=8)
    assert_exact(c.decode_ms_batch(llrs, 50), want, "pageable")
    t_pinned = torch.from_numpy(llrs).pin_memory()
    got = c.decode_ms_batch(t_pinned, 50)
    assert_exact([g.numpy() for g in got], want, "pinned")
    t_dev = torch.from_numpy(llrs).cuda()
    got = c.decode_ms_batch(t_dev, 50)
    torch.cuda.synchronize()
    assert_exact([g.cpu().numpy() for g in got], want, "device")
    # odd offsets / unaligned host views
    raw = np.zeros(llrs.size + 3, np.int8)
    raw[3:] = llrs.ravel()
    got = c.decode_ms_batch(raw[3:].reshape(llrs.shape), 50)
    assert_exact(got, want, "unaligned")
    # mixed pointer kinds are rejected
    with pytest.raises(ldpc.LdpcError):
        c.decode_ms_batch(t_dev, 50, output=np.zeros((700, c.output_len()), np.uint8))


def test_many_chunks_pipeline(oracle):
    """Run in a subprocess with a 1 MiB chunk target so a modest batch crosses many
    pipeline chunks (buffer reuse across the 3 slots)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle"); sys.path.insert(0, %r + "/tests")
import labrador_ldpc_b200 as L, pyoracle
from frames import make_frames
o = pyoracle.Oracle()
code = 2
_, _, llrs = make_frames(o, code, 9001, 3.0, seed=3, ty="i8")
want = o.decode_ms_batch(code, llrs, 20, nthreads=8)
got = L.LDPCCode(code).decode_ms_batch(llrs, 20)
assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want))
print("OK")
''' % (root, root, root)
    env = dict(os.environ, LABRADOR_LDPC_CHUNK_MB="1")
    out = subprocess.check_output([sys.executable, "-c", script], env=env, text=True)
    assert "OK" in out


def test_full_size_properties_tm8192(ldpc, oracle):
    """BASELINE config 3 shape at a size the GPU handles in seconds: properties that need no
    oracle (decoded data == transmitted data on success, re-encoding the decoded data gives
    the decoded codeword, iteration counts in range), plus exact parity on a prefix sample."""
    import torch
    code = 8
    c = ldpc.LDPCCode(code)
    batch = 16384
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    data = torch.randint(0, 256, (batch, c.k() // 8), dtype=torch.uint8, device="cuda", generator=g)
    cw = c.copy_encode_batch(data)
    bits = ((cw.unsqueeze(-1) >> torch.arange(7, -1, -1, device="cuda", dtype=torch.uint8)) & 1).reshape(batch, -1)
    sigma2 = 1.0 / (2.0 * 0.5 * 10.0 ** (2.0 / 10.0))
    y = (1.0 - 2.0 * bits.float()) + (sigma2 ** 0.5) * torch.randn(bits.shape, device="cuda", generator=g)
    llrs = torch.clamp(torch.round(4.0 * 2.0 * y / sigma2), -31, 31).to(torch.int8).contiguous()
    out, ok, iters = c.decode_ms_batch(llrs, 100)
    torch.cuda.synchronize()
    okb = ok.bool()
    assert okb.float().mean().item() > 0.99
    assert torch.equal(out[okb][:, : c.k() // 8], data[okb])
    assert torch.equal(out[okb][:, : c.n() // 8], cw[okb])
    reenc = c.copy_encode_batch(out[:, : c.k() // 8].contiguous())
    assert torch.equal(reenc[okb], out[okb][:, : c.n() // 8])
    assert int(iters[okb].max()) < 100 and bool((iters[~okb] == 100).all())
    # exact parity with the oracle on a prefix of the very same LLR bytes
    ns = 2048
    sample = llrs[:ns].cpu().numpy()
    want = oracle.decode_ms_batch(code, sample, 100, nthreads=16)
    assert_exact((out[:ns].cpu().numpy(), ok[:ns].cpu().numpy(), iters[:ns].cpu().numpy()), want, "prefix sample")


def test_shutdown_and_reinitialise(ldpc, oracle):
    """labrador_ldpc_cuda_shutdown releases every device resource; the next call re-initialises transparently."""
    c = ldpc.LDPCCode.TM2048
    _, _, llrs = make_frames(oracle, 5, 32, 2.0, seed=5, ty="i8")
    want = oracle.decode_ms_batch(5, llrs, 50, nthreads=8)
    assert_exact(c.decode_ms_batch(llrs, 50), want, "before shutdown")
    ldpc.shutdown()
    assert ldpc.lib.labrador_ldpc_cuda_device_count() == 0
    assert_exact(c.decode_ms_batch(llrs, 50), want, "after shutdown (lazy re-init)")
    ldpc.shutdown()
    ldpc.init([0])
    assert ldpc.lib.labrador_ldpc_cuda_device_count() == 1
    _, _, rx = hard_frames(oracle, 5, 16, 3, seed=6)
    assert_exact(c.decode_bf_batch(rx, 20), oracle.decode_bf_batch(5, rx, 20), "bf after explicit re-init")


def test_c_examples_run(tmp_path):
    """The C programs under examples/ (single-codeword reference-signature calls; batched Monte-Carlo trial)
    run against the shared library and report success."""
    import subprocess
    from test_capi_host import build_example
    for name in ("example", "batch_example"):
        res = subprocess.run([build_example(name, tmp_path)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
    assert "frame errors" in res.stdout


def test_launch_counter_and_kernel_names(ldpc):
    before = ldpc.kernel_launch_count()
    c = ldpc.LDPCCode.TC128
    c.decode_ms_batch(np.zeros((4, 128), np.int8), 3)
    assert ldpc.kernel_launch_count() > before
    for ty in ("i8", "i16", "i32", "f32", "f64"):
        assert c.decode_ms_kernel_name(ty).startswith("ms_")


def test_unaligned_device_llrs_take_plain_load_path(ldpc, oracle):
    """The TM kernel stages frames with 16-byte bulk copies; a misaligned device buffer must fall back to
    plain loads and still match the oracle (and odd batch sizes exercise the claim-ahead frame queue)."""
    import torch
    code = 8
    c = ldpc.LDPCCode(code)
    for batch, off in ((1, 0), (3, 5), (149, 0), (301, 1)):
        _, _, llrs = make_frames(oracle, code, batch, 2.0, seed=900 + batch, ty="i8")
        want = oracle.decode_ms_batch(code, llrs, 60, nthreads=8)
        raw = torch.zeros(llrs.size + 16, dtype=torch.int8, device="cuda")
        view = raw[off:off + llrs.size].view(batch, c.n())
        view.copy_(torch.from_numpy(llrs))
        assert view.data_ptr() % 16 == off % 16
        got = c.decode_ms_batch(view, 60)
        torch.cuda.synchronize()
        assert_exact([g.cpu().numpy() for g in got], want, "batch=%d offset=%d" % (batch, off))


def test_generic_kernels_stay_exact():
    """Every (code, type) now has a specialised min-sum kernel; the table-driven generic kernels remain as
    the fallback / A-B reference (LABRADOR_LDPC_FORCE_GENERIC=1) and must stay bit-exact too."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle"); sys.path.insert(0, %r + "/tests")
import labrador_ldpc_b200 as L, pyoracle
from frames import make_frames, hard_frames
o = pyoracle.Oracle()
for code, ty, eb in ((0, "i8", 3.0), (2, "f32", 2.5), (3, "i16", 3.6), (5, "i8", 1.8), (8, "i8", 1.6), (6, "f64", 3.4)):
    c = L.LDPCCode(code)
    assert c.decode_ms_kernel_name(ty).startswith("ms_generic"), c.decode_ms_kernel_name(ty)
    _, _, llrs = make_frames(o, code, 48, eb, seed=31 + code, ty=ty)
    want = o.decode_ms_batch(code, llrs, 40, nthreads=8)
    got = c.decode_ms_batch(llrs, 40)
    assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), (code, ty)
for code in (3, 5, 8):
    _, _, rx = hard_frames(o, code, 32, 5, seed=code)
    want = o.decode_bf_batch(code, rx, 30)
    got = L.LDPCCode(code).decode_bf_batch(rx, 30)
    assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), code
print("OK")
''' % (root, root, root)
    env = dict(os.environ, LABRADOR_LDPC_FORCE_GENERIC="1")
    out = subprocess.check_output([sys.executable, "-c", script], env=env, text=True)
    assert "OK" in out


def test_specialised_kernel_dispatch(ldpc):
    names = {(code, ty): ldpc.LDPCCode(code).decode_ms_kernel_name(ty) for code in range(9)
             for ty in ("i8", "i16", "i32", "f32", "f64")}
    if os.environ.get("LABRADOR_LDPC_FORCE_GENERIC") == "1":
        pytest.skip("generic forced")
    for (code, ty), name in names.items():
        if code < 3:
            assert name == ("ms_tc_x2<i8>" if ty == "i8" else "ms_tc_warp<%s>" % ty)
        elif ty == "i8":
            assert name == "ms_tm_s16x2<i8>"                # TM1280 too since round 2 (M = 128 on the packed kernel)
        elif ty == "i16":
            assert name == "ms_tm_s16x2<i16>"               # packed 16-bit check side (decode_ms_tm_i16.cu)
        else:
            assert name == "ms_tm_wide<%s>" % ty
    for code in (9, 10, 11):                                 # the k = 16384 codes: one codeword per cluster of four CTAs (i8)
        assert ldpc.LDPCCode(code).decode_ms_kernel_name("i8") == "ms_tm_cluster<i8>"
        assert ldpc.LDPCCode(code).decode_ms_kernel_name("f32") == "ms_generic<f32>"


def test_mixed_code_batch_on_concurrent_streams(ldpc, oracle):
    """BASELINE config 4: a high-rate mixed batch (TM5120 + TM6144) as two homogeneous sub-batches on
    concurrent streams; each must match the oracle."""
    import torch
    jobs, wants = [], []
    for code, eb in ((6, 3.4), (7, 2.4), (6, 4.0)):
        _, _, llrs = make_frames(oracle, code, 96, eb, seed=70 + code, ty="i8")
        wants.append(oracle.decode_ms_batch(code, llrs, 60, nthreads=8))
        jobs.append((ldpc.LDPCCode(code), torch.from_numpy(llrs).cuda()))
    results = ldpc.decode_ms_mixed(jobs, 60)
    torch.cuda.synchronize()
    for got, want, (code, _) in zip(results, wants, jobs):
        assert_exact([g.cpu().numpy() for g in got], want, "mixed batch %s" % code.name)


def test_in_process_multi_device_split(oracle):
    """labrador_ldpc_cuda_init(devices) with several GPUs: a host-pointer batch is split into contiguous
    per-device shards (one host thread per device, no collective).  Needs >= 2 GPUs; runs in a subprocess
    so the device list of this process stays untouched."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle"); sys.path.insert(0, %r + "/tests")
import labrador_ldpc_b200 as L, pyoracle
from frames import make_frames
o = pyoracle.Oracle()
L.init([0, 1])
assert L.lib.labrador_ldpc_cuda_device_count() == 2
for code, batch in ((8, 301), (0, 1001)):
    _, cw, llrs = make_frames(o, code, batch, 2.5, seed=5, ty="i8")
    want = o.decode_ms_batch(code, llrs, 40, nthreads=8)
    got = L.LDPCCode(code).decode_ms_batch(llrs, 40)
    assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), code
    data = cw[:, : o.k(code) // 8].copy()
    assert np.array_equal(L.LDPCCode(code).copy_encode_batch(data), cw)
print("OK")
''' % (root, root, root)
    out = subprocess.check_output([sys.executable, "-c", script], text=True)
    assert "OK" in out


def test_cuda_tensor_calls_follow_the_current_stream(ldpc, oracle):
    """Single-codeword calls on CUDA tensors are ordered on torch's CURRENT stream (side streams are non-blocking
    with respect to the legacy default stream): input produced on a side stream right before the call must be seen."""
    import torch
    code = 5
    c = ldpc.LDPCCode(code)
    _, cw, llrs = make_frames(oracle, code, 1, 3.0, seed=77, ty="i8")
    want_ok, want_it, want_out = oracle.decode_ms(code, llrs[0], 50)
    src = torch.from_numpy(llrs[0]).cuda()
    data = torch.from_numpy(cw[0, : c.k() // 8].copy()).cuda()
    busy = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(8):
            busy.add_(1)                       # a few ms of work ahead of the producer on the side stream
        llr_d = torch.zeros(c.n(), dtype=torch.int8, device="cuda")
        llr_d.copy_(src)                       # the producer
        out_d = torch.zeros(c.output_len(), dtype=torch.uint8, device="cuda")
        ok, it = c.decode_ms(llr_d, out_d, maxiters=50)
        cw_d = torch.zeros(c.n() // 8, dtype=torch.uint8, device="cuda")
        dat_d = torch.zeros(c.k() // 8, dtype=torch.uint8, device="cuda")
        dat_d.copy_(data)
        c.copy_encode(dat_d, cw_d)
        h_d = torch.zeros(c.n() // 8, dtype=torch.uint8, device="cuda")
        c.llrs_to_hard(llr_d, h_d)
    side.synchronize()
    assert (ok, it) == (want_ok, want_it)
    assert np.array_equal(out_d.cpu().numpy(), want_out)
    assert np.array_equal(cw_d.cpu().numpy(), cw[0])
    assert np.array_equal(h_d.cpu().numpy(), oracle.llrs_to_hard(code, llrs[0]))


def test_many_async_launches_on_several_streams(ldpc, oracle):
    """More stream-ordered launches in flight than the context has work-counter slots (64), spread over four
    streams and three kernel families: every launch must still claim exactly its own frames."""
    import torch
    jobs = []
    for code, eb, batch in ((4, 2.6, 40), (0, 3.0, 300), (5, 1.8, 24)):
        _, _, llrs = make_frames(oracle, code, batch, eb, seed=900 + code, ty="i8")
        jobs.append((ldpc.LDPCCode(code), torch.from_numpy(llrs).cuda(), oracle.decode_ms_batch(code, llrs, 30, nthreads=8)))
    streams = [torch.cuda.Stream() for _ in range(4)]
    results = []
    torch.cuda.synchronize()
    for i in range(200):
        c, l, want = jobs[i % len(jobs)]
        st = streams[i % len(streams)]
        with torch.cuda.stream(st):
            results.append((c.decode_ms_batch(l, 30, stream=st.cuda_stream), want, c.name))
    torch.cuda.synchronize()
    for got, want, name in results:
        assert_exact([g.cpu().numpy() for g in got], want, "async " + name)


@pytest.mark.parametrize("env", [{"LABRADOR_LDPC_TC_X2_HABS": "0"}, {"LABRADOR_LDPC_TC_X2_HABS": "1"},
                                 {"LABRADOR_LDPC_TC_X2_HABS": "2"}, {"LABRADOR_LDPC_TC_X2_HABS": "3"},
                                 {"LABRADOR_LDPC_TM_ARITH": "632"}, {"LABRADOR_LDPC_TM_ARITH": "932"},
                                 {"LABRADOR_LDPC_TM_ARITH": "5"}, {"LABRADOR_LDPC_TM_ARITH": "1132"}, {"LABRADOR_LDPC_TM_WPT": "2"}],
                         ids=["tc-int", "tc-fp16-minima", "tc-fp16-check-side", "tc-hard-bits-in-messages",
                              "tm-arith-632", "tm-arith-932", "tm-arith-5", "tm-arith-1132", "tm8192-two-slots-per-thread"])
def test_i8_min_sum_arithmetic_variants_stay_exact(env):
    """The i8 min-sum kernels keep their earlier arithmetic forms selectable (integer lanes, fp16 only for the minima,
    fp16 check side, hard decisions inside the messages): every one of them must reproduce the oracle bit for bit,
    including frames that saturate and frames that run into the iteration cap."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    codes = (0, 1, 2) if "LABRADOR_LDPC_TC_X2_HABS" in env else (3, 4, 5, 6, 7, 8)
    script = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle"); sys.path.insert(0, %r + "/tests")
import labrador_ldpc_b200 as L, pyoracle
from frames import make_frames
o = pyoracle.Oracle()
EB = {0: 2.0, 1: 2.0, 2: 1.8, 3: 3.4, 4: 2.4, 5: 1.6, 6: 3.4, 7: 2.4, 8: 1.6}
for code in %r:
    c = L.LDPCCode(code)
    _, _, llrs = make_frames(o, code, 203, EB[code], seed=77 + code, ty="i8")
    rng = np.random.default_rng(code)
    llrs[:8] = rng.integers(-128, 128, llrs[:8].shape).astype(np.int8)        # saturation stress, never converges
    llrs[8, :] = -128
    for mi in (25, 2):
        want = o.decode_ms_batch(code, llrs, mi, nthreads=16)
        got = c.decode_ms_batch(llrs, mi)
        assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), (code, mi)
print("OK")
''' % (root, root, root, codes)
    out = subprocess.check_output([sys.executable, "-c", script], env=dict(os.environ, LABRADOR_LDPC_NO_REBUILD="1", **env), text=True)
    assert "OK" in out


@pytest.mark.parametrize("code", list(range(9)))
def test_i8_unstructured_llrs_run_the_whole_iteration_budget(ldpc, oracle, code):
    """LLRs that are not a noisy codeword at all -- uniform over the whole i8 range, sparse large values among zeros,
    two-valued +-1, long runs of -128 -- never converge: 100 iterations of saturating additions, sign flips on almost
    every edge (the self-correction rule fires constantly) and ties in the minima.  Decoded bytes, flags and iteration
    counts must equal the oracle's."""
    c = ldpc.LDPCCode(code)
    n = c.n()
    rng = np.random.default_rng(4242 + code)
    rows = [rng.integers(-128, 128, (24, n)),                                   # uniform
            rng.integers(-128, 128, (8, n)) * (rng.random((8, n)) < 0.1),       # sparse among zeros
            rng.choice([-1, 1], (8, n)),                                        # smallest magnitudes: ties everywhere
            rng.choice([-128, -127, 127], (8, n)),                              # the three extreme values
            np.where(rng.random((8, n)) < 0.5, -128, rng.integers(-3, 4, (8, n)))]
    llrs = np.concatenate(rows).astype(np.int8)
    for mi in (100, 7):
        want = oracle.decode_ms_batch(code, llrs, mi, nthreads=16)
        got = c.decode_ms_batch(llrs, mi)
        assert_exact(got, want, "%s unstructured LLRs, maxiters %d" % (c.name if hasattr(c, "name") else code, mi))
    assert int(np.asarray(want[1]).sum()) < len(llrs)                           # (the cap was really reached)
