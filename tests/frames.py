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


def quantise_soft(soft, scale, limit, dtype):
    """Restatement of the library's soft front end (csrc/front.cuh): everything in float32,
    product NaN -> 0, round half to even, clamp to +-limit."""
    with np.errstate(invalid="ignore", over="ignore"):
        p = soft.astype(np.float32) * np.float32(scale)
    p = np.where(np.isnan(p), np.float32(0), p)
    return np.clip(np.rint(p), -limit, limit).astype(dtype)


def soft_frames(oracle, code, batch, ebn0_db, seed):
    """float32 channel LLRs with the awkward values sprinkled in: NaN, +-inf, exact .5 ties, huge magnitudes."""
    _, cw, llr = make_frames(oracle, code, batch, ebn0_db, seed, ty="f32")
    rng = np.random.default_rng(seed + 7)
    flat = llr.reshape(-1)
    idx = rng.choice(flat.size, size=min(flat.size // 50, 4000), replace=False)
    specials = np.array([np.nan, np.inf, -np.inf, 0.125, -0.125, 0.375, -0.375, 0.625, 1e30, -1e30, 0.0, -0.0,
                         7.875, -7.875, 31.5 / 4, 1e-40], dtype=np.float32)
    flat[idx] = specials[rng.integers(0, specials.size, idx.size)]
    return cw, llr


def philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): ctr [..., 4] and key [..., 2] uint32 -> [..., 4] uint32.
    Test-side restatement of the generator in csrc/channel.cu; pinned by the Random123 known answers in
    tests/test_capi_host.py."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k = [np.asarray(key[..., i], dtype=np.uint64) for i in range(2)]
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & m32, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & m32]
        k = [(k[0] + np.uint64(0x9E3779B9)) & m32, (k[1] + np.uint64(0xBB67AE85)) & m32]
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_words(seed, frames, words, stream):
    """Philox output for counters (frame, j, stream), j < words: uint32 [len(frames), words, 4]."""
    frames = np.asarray(frames, dtype=np.uint64)
    ctr = np.zeros((len(frames), words, 4), dtype=np.uint32)
    ctr[..., 0] = (frames & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
    ctr[..., 1] = (frames >> np.uint64(32)).astype(np.uint32)[:, None]
    ctr[..., 2] = np.arange(words, dtype=np.uint32)[None, :]
    ctr[..., 3] = stream
    key = np.zeros((len(frames), words, 2), dtype=np.uint32)
    key[..., 0] = seed & 0xFFFFFFFF
    key[..., 1] = (seed >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr, key)
