"""GPU parity tests of the fused front ends (include/labrador_ldpc.h "Fused front ends").

decode_ms_{i8,i16}_soft_batch and decode_ms_i8_hard_batch must give exactly the result of the two-step
sequence the reference's callers run -- quantise / hard_to_llrs (reference src/decoder.rs:484-493), then
decode_ms::<i8|i16> (:347-475) -- here computed by the CPU oracle on the same inputs.
"""
import os

import numpy as np
import pytest

from frames import hard_frames, quantise_soft, soft_frames
from test_gpu_parity import CODES, EBN0, NAMES, assert_exact

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("code", CODES)
def test_soft_front_i8_exact(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    batch = 256 if code < 6 else 96
    _, soft = soft_frames(oracle, code, batch, EBN0[code], seed=900 + code)
    for scale, limit in ((4.0, 31), (24.0, 127)):
        q = quantise_soft(soft, scale, limit, np.int8)
        assert np.array_equal(c.quantise_batch(soft, scale, limit, "i8"), q), "stand-alone quantiser"
        want = oracle.decode_ms_batch(code, q, 60, nthreads=8)
        got = c.decode_ms_soft_batch(soft, scale, limit, 60, "i8")
        assert_exact(got, want, "%s soft->i8 scale=%g limit=%d" % (NAMES[code], scale, limit))
        assert 0 < want[1].sum()


@pytest.mark.parametrize("code", CODES)
def test_soft_front_i16_exact(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    batch = 96 if code < 6 else 40
    _, soft = soft_frames(oracle, code, batch, EBN0[code], seed=950 + code)
    q = quantise_soft(soft, 256.0, 8191, np.int16)
    assert np.array_equal(c.quantise_batch(soft, 256.0, 8191, "i16"), q)
    want = oracle.decode_ms_batch(code, q, 50, nthreads=8)
    got = c.decode_ms_soft_batch(soft, 256.0, 8191, 50, "i16")
    assert_exact(got, want, "%s soft->i16" % NAMES[code])


@pytest.mark.parametrize("code", CODES)
def test_hard_front_exact(ldpc, oracle, code):
    c = ldpc.LDPCCode(code)
    batch = 128 if code < 6 else 48
    for flips in (0, 2, 9):
        _, _, rx = hard_frames(oracle, code, batch, flips, seed=40 + code + flips)
        llrs = np.stack([oracle.hard_to_llrs(code, r, "i8") for r in rx])
        want = oracle.decode_ms_batch(code, llrs, 50, nthreads=8)
        got = c.decode_ms_hard_batch(rx, 50)
        assert_exact(got, want, "%s hard front, %d flips" % (NAMES[code], flips))
        # and it equals the library's own two-step path
        two = c.decode_ms_batch(c.hard_to_llrs_batch(rx, "i8"), 50)
        assert_exact(got, two, "%s hard front vs two-step" % NAMES[code])


def test_front_device_pointers_and_unaligned(ldpc, oracle):
    """Device-resident soft values (stream-ordered entry point), including a view whose base is not
    16-byte aligned (the kernel then skips the bulk copy and loads directly)."""
    import torch
    for code in (2, 5, 8):
        c = ldpc.LDPCCode(code)
        _, soft = soft_frames(oracle, code, 40, EBN0[code], seed=970 + code)
        q = quantise_soft(soft, 4.0, 31, np.int8)
        want = oracle.decode_ms_batch(code, q, 50, nthreads=8)
        for off in (0, 1):
            buf = torch.zeros(soft.size + 4, dtype=torch.float32, device="cuda")
            view = buf[off: off + soft.size].view(soft.shape)
            view.copy_(torch.from_numpy(soft))
            got = c.decode_ms_soft_batch(view, 4.0, 31, 50, "i8")
            torch.cuda.synchronize()
            assert_exact([g.cpu().numpy() for g in got], want, "%s device soft off=%d" % (NAMES[code], off))
        _, _, rx = hard_frames(oracle, code, 24, 3, seed=code)
        llrs = np.stack([oracle.hard_to_llrs(code, r, "i8") for r in rx])
        want = oracle.decode_ms_batch(code, llrs, 50, nthreads=8)
        hb = torch.zeros(rx.size + 16, dtype=torch.uint8, device="cuda")
        for off in (0, 3):
            view = hb[off: off + rx.size].view(rx.shape)
            view.copy_(torch.from_numpy(rx))
            got = c.decode_ms_hard_batch(view, 50)
            torch.cuda.synchronize()
            assert_exact([g.cpu().numpy() for g in got], want, "%s device hard off=%d" % (NAMES[code], off))


def test_front_generic_kernels_stay_exact():
    """The table-driven generic kernels carry the same front ends (LABRADOR_LDPC_FORCE_GENERIC=1)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle"); sys.path.insert(0, %r + "/tests")
import labrador_ldpc_b200 as L, pyoracle
from frames import soft_frames, hard_frames, quantise_soft
o = pyoracle.Oracle()
same = lambda got, want: all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want))
for code, eb in ((1, 3.0), (3, 3.6), (6, 3.4), (8, 1.6)):
    c = L.LDPCCode(code)
    assert c.decode_ms_kernel_name("i8").startswith("ms_generic")
    _, soft = soft_frames(o, code, 32, eb, seed=77 + code)
    for ty, scale, limit, dt in (("i8", 4.0, 31, np.int8), ("i16", 256.0, 8191, np.int16)):
        want = o.decode_ms_batch(code, quantise_soft(soft, scale, limit, dt), 40, nthreads=8)
        assert same(c.decode_ms_soft_batch(soft, scale, limit, 40, ty), want), (code, ty)
    _, _, rx = hard_frames(o, code, 24, 4, seed=code)
    want = o.decode_ms_batch(code, np.stack([o.hard_to_llrs(code, r, "i8") for r in rx]), 40, nthreads=8)
    assert same(c.decode_ms_hard_batch(rx, 40), want), code
print("OK")
''' % (root, root, root)
    env = dict(os.environ, LABRADOR_LDPC_FORCE_GENERIC="1")
    out = subprocess.check_output([sys.executable, "-c", script], env=env, text=True)
    assert "OK" in out
