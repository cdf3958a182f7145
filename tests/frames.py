"""Seeded synthetic frames for the parity tests (SURVEY.md section 8d channel model).

random data -> reference-exact encode (oracle) -> BPSK x = 1-2b -> y = x + sigma*z,
sigma^2 = 1/(2*(k/n)*10^(EbN0/10)) -> LLR = 2y/sigma^2
  i8 : clamp(round(4*LLR),  -31,   31)
  i16: clamp(round(256*LLR), -8191, 8191)
  i32: round(65536*LLR)
  f32/f64: LLR as is
"""
import numpy as np


def make_frames(oracle, code, batch, ebn0_db, seed, ty="i8"):
    rng = np.random.default_rng(seed)
    n, k = oracle.n(code), oracle.k(code)
    data = rng.integers(0, 256, (batch, k // 8), dtype=np.uint8)
    cw = oracle.copy_encode_batch(code, data, nthreads=4)
    bits = np.unpackbits(cw, axis=1).astype(np.float64)
    sigma2 = 1.0 / (2.0 * (k / n) * 10.0 ** (ebn0_db / 10.0))
    y = (1.0 - 2.0 * bits) + np.sqrt(sigma2) * rng.standard_normal(bits.shape)
    llr = 2.0 * y / sigma2
    if ty == "i8":
        q = np.clip(np.rint(4.0 * llr), -31, 31).astype(np.int8)
    elif ty == "i16":
        q = np.clip(np.rint(256.0 * llr), -8191, 8191).astype(np.int16)
    elif ty == "i32":
        q = np.rint(65536.0 * llr).astype(np.int32)
    elif ty == "f32":
        q = llr.astype(np.float32)
    elif ty == "f64":
        q = llr.astype(np.float64)
    else:
        raise ValueError(ty)
    return data, cw, np.ascontiguousarray(q)


def hard_frames(oracle, code, batch, flips, seed):
    """Codewords with `flips` random bit errors each (for decode_bf)."""
    rng = np.random.default_rng(seed)
    n, k = oracle.n(code), oracle.k(code)
    data = rng.integers(0, 256, (batch, k // 8), dtype=np.uint8)
    cw = oracle.copy_encode_batch(code, data, nthreads=4)
    rx = cw.copy()
    for f in range(batch):
        pos = rng.choice(n, size=flips, replace=False)
        for p in pos:
            rx[f, p // 8] ^= 1 << (7 - (p % 8))
    return data, cw, rx
