"""The three k = 16384 TM codes (TM20480 r=4/5, TM24576 r=2/3, TM32768 r=1/2).

The reference ships their parity-check constants (src/codes/compact_parity_checks.rs:84-96, PHI_J_K_M4096 / M8192,
selected at src/codes/mod.rs:473-476) but neither parameters nor generators (src/lib.rs:81-83), so NO reference golden
exists for them: parity here is against the oracle only -- the reference's decoder algorithms, restated, run over
these codes' edge lists -- and the oracle's edge lists are pinned the only way available: codewords produced by the
product's sparse-H encoder must satisfy every one of the oracle's parity checks (tests/test_capi_host.py does the same
on the host tables, without a GPU).
"""
import numpy as np
import pytest

from test_gpu_parity import assert_exact, assert_float_parity

pytestmark = pytest.mark.gpu

CODES = {9: "TM20480", 10: "TM24576", 11: "TM32768"}
EBN0 = {9: 3.3, 10: 2.2, 11: 1.4}          # mostly, not always, decodable: a mix of iteration counts


def gpu_codewords(ldpc, code, batch, seed):
    c = ldpc.LDPCCode(code)
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, (batch, c.k() // 8), dtype=np.uint8)
    data[0, :] = 0
    data[1, :] = 0xFF
    cw = c.copy_encode_batch(data)
    assert cw.shape == (batch, c.n() // 8) and np.array_equal(cw[:, : c.k() // 8], data)
    return data, cw


def syndrome_weight(oracle, code, full_bits):
    """Number of unsatisfied checks of the ORACLE's edge list for [batch, n+p] bit arrays."""
    _, checks, vars_, _ = oracle.edges(code)
    nc = oracle.n(code) + oracle.p(code) - oracle.k(code)
    out = []
    for bits in full_bits:
        out.append(int((np.bincount(checks, weights=bits[vars_], minlength=nc).astype(np.int64) & 1).sum()))
    return out


def frames(ldpc, code, batch, ebn0, seed, ty):
    c = ldpc.LDPCCode(code)
    _, cw = gpu_codewords(ldpc, code, batch, seed)
    rng = np.random.default_rng(seed + 1)
    bits = np.unpackbits(cw, axis=1).astype(np.float64)
    sigma2 = 1.0 / (2.0 * (c.k() / c.n()) * 10.0 ** (ebn0 / 10.0))
    llr = 2.0 * ((1.0 - 2.0 * bits) + np.sqrt(sigma2) * rng.standard_normal(bits.shape)) / sigma2
    if ty == "i8":
        return cw, np.clip(np.rint(4.0 * llr), -31, 31).astype(np.int8)
    if ty == "i16":
        return cw, np.clip(np.rint(256.0 * llr), -8191, 8191).astype(np.int16)
    return cw, llr.astype({"f32": np.float32, "f64": np.float64}[ty])


@pytest.mark.parametrize("code", list(CODES))
def test_sparse_h_encoder_satisfies_the_oracles_checks(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    _, cw = gpu_codewords(ldpc, code, 37, seed=code)
    assert not cw[0].any()                                              # all-zero data -> all-zero codeword
    # decoding the clean codeword reproduces it AND yields the punctured bits: all n+p bits must satisfy H
    llrs = c.hard_to_llrs_batch(cw, "i8")
    out, ok, iters = c.decode_ms_batch(llrs, 20)
    assert ok.all() and np.array_equal(out[:, : c.n() // 8], cw)
    assert syndrome_weight(oracle, code, np.unpackbits(out, axis=1)) == [0] * len(cw)
    want = oracle.decode_ms_batch(code, llrs, 20, nthreads=16)
    assert_exact((out, ok, iters), want, CODES[code] + " clean codewords")
    # in place == copy
    buf = np.zeros_like(cw)
    buf[:, : c.k() // 8] = cw[:, : c.k() // 8]
    assert np.array_equal(c.encode_batch(buf), cw)


@pytest.mark.parametrize("code", list(CODES))
@pytest.mark.parametrize("ty", ["i8", "i16", "f32"])
def test_decode_ms_awgn_matches_oracle(ldpc, oracle, code, ty):
    c = ldpc.LDPCCode(code)
    cw, llrs = frames(ldpc, code, 24, EBN0[code], seed=300 + code, ty=ty)
    want = oracle.decode_ms_batch(code, llrs, 60, nthreads=16)
    got = c.decode_ms_batch(llrs, 60)
    if ty in ("i8", "i16"):
        assert_exact(got, want, "%s %s" % (CODES[code], ty))
    else:
        assert_float_parity(got, want, "%s %s" % (CODES[code], ty), min_frac=1.0)
    okw = want[1].astype(bool)
    assert okw.any() and np.array_equal(want[0][okw][:, : c.n() // 8], cw[okw])


@pytest.mark.parametrize("code", list(CODES))
def test_decode_ms_saturation_stress(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    rng = np.random.default_rng(500 + code)
    llrs = rng.integers(-128, 128, (6, c.n())).astype(np.int8)
    llrs[0, :] = -128
    llrs[1, ::2] = 127
    for maxiters in (1, 3):
        assert_exact(c.decode_ms_batch(llrs, maxiters), oracle.decode_ms_batch(code, llrs, maxiters, nthreads=8),
                     "%s stress %d" % (CODES[code], maxiters))


@pytest.mark.parametrize("code", list(CODES))
def test_decode_bf_and_erasures_match_oracle(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    _, cw = gpu_codewords(ldpc, code, 48, seed=700 + code)
    rng = np.random.default_rng(701 + code)
    rx = cw.copy()
    for f in range(len(rx)):
        for p in rng.choice(c.n(), size=f % 9, replace=False):
            rx[f, p // 8] ^= 1 << (7 - (p % 8))
    want = oracle.decode_bf_batch(code, rx, 40, nthreads=16)
    got = c.decode_bf_batch(rx, 40)
    assert_exact(got, want, CODES[code] + " bf")
    assert want[1][0] and np.array_equal(want[0][0][: c.n() // 8], cw[0])
    # the table-driven kernel agrees too (it is the fallback for these codes' generic paths)
    import os, subprocess, sys
    env = dict(os.environ, LABRADOR_LDPC_FORCE_GENERIC="1", LABRADOR_LDPC_NO_REBUILD="1")
    script = ("import sys, numpy as np; sys.path.insert(0, %r); import labrador_ldpc_b200 as L; "
              "rx = np.load(sys.argv[1]); out, ok, it = L.LDPCCode(%d).decode_bf_batch(rx, 40); "
              "np.savez(sys.argv[2], out=out, ok=ok, it=it)" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), code))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        np.save(d + "/rx.npy", rx)
        subprocess.check_call([sys.executable, "-c", script, d + "/rx.npy", d + "/o.npz"], env=env)
        z = np.load(d + "/o.npz")
        assert_exact((z["out"], z["ok"], z["it"]), want, CODES[code] + " bf generic")


def test_converters_and_reference_signature_calls(ldpc, oracle):
    code = 9
    c = ldpc.LDPCCode(code)
    _, cw = gpu_codewords(ldpc, code, 3, seed=9)
    for ty in ("i8", "f32"):
        llrs = c.hard_to_llrs_batch(cw, ty)
        assert np.array_equal(llrs[1], oracle.hard_to_llrs(code, cw[1], ty))
        assert np.array_equal(c.llrs_to_hard_batch(llrs), cw)
    out = np.zeros(c.output_len(), np.uint8)
    llr1 = c.hard_to_llrs_batch(cw[2:3], "i8")[0]
    ok, it = c.decode_ms(llr1, out, maxiters=10)
    wok, wit, wout = oracle.decode_ms(code, llr1, 10)
    assert (ok, it) == (wok, wit) and np.array_equal(out, wout)


@pytest.mark.parametrize("code", list(CODES))
def test_cluster_kernel_larger_batch_and_unaligned_buffers(ldpc, oracle, code):
    """More frames than resident clusters (every cluster decodes several frames: staging double buffer, state
    re-initialisation, static frame assignment), tight iteration caps, and a misaligned LLR buffer (plain-load path)."""
    import torch
    c = ldpc.LDPCCode(code)
    cw, llrs = frames(ldpc, code, 150, EBN0[code] + 0.3, seed=800 + code, ty="i8")
    for maxiters in (100, 7, 1):
        want = oracle.decode_ms_batch(code, llrs, maxiters, nthreads=16)
        got = c.decode_ms_batch(llrs, maxiters)
        assert_exact(got, want, "%s cluster kernel maxiters=%d" % (CODES[code], maxiters))
    raw = torch.zeros(llrs.size + 64, dtype=torch.int8, device="cuda")
    for off in (1, 8):
        view = raw[off:off + llrs.size].view(llrs.shape)
        view.copy_(torch.from_numpy(llrs))
        got = c.decode_ms_batch(view, 30)
        torch.cuda.synchronize()
        assert_exact([g.cpu().numpy() for g in got], oracle.decode_ms_batch(code, llrs, 30, nthreads=16),
                     "%s cluster kernel, LLRs at byte offset %d" % (CODES[code], off))
    # the table-driven kernel (the path of the other LLR types) gives the same answer
    import os, subprocess, sys, tempfile
    env = dict(os.environ, LABRADOR_LDPC_TM_CLUSTER="0", LABRADOR_LDPC_NO_REBUILD="1")
    script = ("import sys, numpy as np; sys.path.insert(0, %r); import labrador_ldpc_b200 as L; "
              "x = np.load(sys.argv[1]); out, ok, it = L.LDPCCode(%d).decode_ms_batch(x, 30); "
              "np.savez(sys.argv[2], out=out, ok=ok, it=it)" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), code))
    with tempfile.TemporaryDirectory() as d:
        np.save(d + "/x.npy", llrs[:40])
        subprocess.check_call([sys.executable, "-c", script, d + "/x.npy", d + "/o.npz"], env=env)
        z = np.load(d + "/o.npz")
        assert_exact((z["out"], z["ok"], z["it"]), oracle.decode_ms_batch(code, llrs[:40], 30, nthreads=16), CODES[code] + " table-driven")


@pytest.mark.parametrize("env", [{"LABRADOR_LDPC_CLUSTER_ASYNC": "0"}, {"LABRADOR_LDPC_CLUSTER_QUAD": "0"},
                                 {"LABRADOR_LDPC_CLUSTER_QUAD": "0", "LABRADOR_LDPC_CLUSTER_ASYNC": "1"},
                                 {"LABRADOR_LDPC_CLUSTER_MINB": "1"}],
                         ids=["quad-barriers", "pair-barriers", "pair-async", "quad-async-1cta"])
def test_cluster_kernel_variants_match_oracle(ldpc, oracle, env):
    """The shipped kernel packs four messages per word and synchronises through st.async + mbarrier; the earlier forms
    (two messages per word, two cluster barriers per iteration) stay selectable for A/B runs and must stay exact."""
    import os, subprocess, sys, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for code in CODES:
        _, llrs = frames(ldpc, code, 90, EBN0[code], seed=900 + code, ty="i8")
        script = ("import sys, numpy as np; sys.path.insert(0, %r); import labrador_ldpc_b200 as L; "
                  "x = np.load(sys.argv[1]); out, ok, it = L.LDPCCode(%d).decode_ms_batch(x, 40); "
                  "np.savez(sys.argv[2], out=out, ok=ok, it=it)" % (root, code))
        with tempfile.TemporaryDirectory() as d:
            np.save(d + "/x.npy", llrs)
            subprocess.check_call([sys.executable, "-c", script, d + "/x.npy", d + "/o.npz"],
                                  env=dict(os.environ, LABRADOR_LDPC_NO_REBUILD="1", **env))
            z = np.load(d + "/o.npz")
            assert_exact((z["out"], z["ok"], z["it"]), oracle.decode_ms_batch(code, llrs, 40, nthreads=16),
                         "%s cluster kernel %r" % (CODES[code], env))


@pytest.mark.parametrize("code", list(CODES))
def test_cluster_kernel_unstructured_llrs(ldpc, oracle, code):
    """LLRs that never converge (uniform over the i8 range, +-1 only, the extreme values): every iteration saturates and
    the self-correction rule fires on most edges; 30 iterations through the cluster kernel against the oracle."""
    c = ldpc.LDPCCode(code)
    n = c.n()
    rng = np.random.default_rng(5151 + code)
    llrs = np.concatenate([rng.integers(-128, 128, (6, n)), rng.choice([-1, 1], (3, n)),
                           rng.choice([-128, -127, 127], (3, n))]).astype(np.int8)
    for mi in (30, 4):
        assert_exact(c.decode_ms_batch(llrs, mi), oracle.decode_ms_batch(code, llrs, mi, nthreads=16),
                     "%s unstructured LLRs, maxiters %d" % (CODES[code], mi))
