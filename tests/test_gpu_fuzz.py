"""Randomised parity: code, LLR type, batch size, iteration cap, buffer placement (numpy / pinned / device, with a byte
offset that breaks the 16-byte alignment the TMA staging needs) drawn at random, every result compared with the oracle.
Covers the kernel families behind one entry point in combinations the targeted tests do not enumerate."""
import numpy as np
import pytest

from frames import hard_frames, make_frames
from test_gpu_parity import EBN0, NAMES, assert_exact, assert_float_parity

pytestmark = pytest.mark.gpu


def place(torch, arr, how, offset):
    """numpy array -> the buffer kind under test (same bytes)."""
    if how == "numpy":
        return arr
    flat = torch.from_numpy(arr.reshape(-1).view(np.uint8).copy())
    if how == "pinned":
        return torch.from_numpy(arr.copy()).pin_memory()
    raw = torch.zeros(flat.numel() + 64, dtype=torch.uint8, device="cuda")
    raw[offset:offset + flat.numel()].copy_(flat)
    tdt = {np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32,
           np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8}[arr.dtype]
    return raw[offset:offset + flat.numel()].view(tdt).view(arr.shape)


def to_np(x):
    return x if isinstance(x, np.ndarray) else x.cpu().numpy()


@pytest.mark.parametrize("seed", range(24))
def test_decode_ms_random_configurations(ldpc, oracle, seed):
    import torch
    rng = np.random.default_rng(9000 + seed)
    code = int(rng.integers(0, 9))
    ty = ["i8", "i16", "i32", "f32", "f64"][int(rng.integers(0, 5))]
    batch = int(rng.choice([1, 2, 3, 7, 33, 64, 149, 300])) if code < 6 else int(rng.choice([1, 2, 5, 37, 80]))
    maxiters = int(rng.choice([0, 1, 2, 5, 13, 40]))
    ebn0 = EBN0[code] + float(rng.choice([-0.6, 0.0, 0.8]))
    how = ["numpy", "pinned", "device"][int(rng.integers(0, 3))]
    size = {"i8": 1, "i16": 2, "i32": 4, "f32": 4, "f64": 8}[ty]
    offset = int(rng.choice([0, size, 16])) if how == "device" else 0
    _, _, llrs = make_frames(oracle, code, batch, ebn0, seed=9100 + seed, ty=ty)
    if ty == "i8" and rng.random() < 0.3:
        llrs = np.clip(llrs.astype(np.int32) * 5, -128, 127).astype(np.int8)       # saturating magnitudes
    want = oracle.decode_ms_batch(code, llrs, maxiters, nthreads=16)
    got = ldpc.LDPCCode(code).decode_ms_batch(place(torch, llrs, how, offset), maxiters)
    torch.cuda.synchronize()
    what = "%s %s batch %d maxiters %d %s+%d" % (NAMES[code], ty, batch, maxiters, how, offset)
    if ty in ("i8", "i16", "i32"):
        assert_exact([to_np(g) for g in got], want, what)
    else:
        assert_float_parity([to_np(g) for g in got], want, what, min_frac=1.0)


@pytest.mark.parametrize("seed", range(10))
def test_decode_bf_and_encode_random_configurations(ldpc, oracle, seed):
    import torch
    rng = np.random.default_rng(9500 + seed)
    code = int(rng.integers(0, 9))
    batch = int(rng.choice([1, 3, 31, 130, 257]))
    flips = int(rng.integers(0, 7))
    how = ["numpy", "pinned", "device"][int(rng.integers(0, 3))]
    offset = int(rng.choice([0, 1, 4])) if how == "device" else 0
    data, cw, rx = hard_frames(oracle, code, batch, flips, seed=9600 + seed)
    c = ldpc.LDPCCode(code)
    maxiters = int(rng.choice([0, 1, 4, 30]))
    got = c.decode_bf_batch(place(torch, rx, how, offset), maxiters)
    torch.cuda.synchronize()
    assert_exact([to_np(g) for g in got], oracle.decode_bf_batch(code, rx, maxiters, nthreads=8),
                 "%s bf batch %d flips %d maxiters %d %s+%d" % (NAMES[code], batch, flips, maxiters, how, offset))
    enc = c.copy_encode_batch(place(torch, data, how, offset))
    torch.cuda.synchronize()
    assert np.array_equal(to_np(enc), cw)
