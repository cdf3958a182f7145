"""The fp16 formulas of the i8 min-sum check side (csrc/decode_ms_tm.cu ARITH 10, decode_ms_tc_x2.cu MODE 2,
decode_ms_tm_cluster.cu) modelled in numpy: every operand combination the kernels can meet gives the integer result
of the reference's rule (src/decoder.rs:398-405, :422-426).  A fused multiply-add is modelled as the exact float64
result rounded once to fp16 -- all products and sums here are far inside float64's exact range.  CPU only; the GPU
parity tests check the kernels themselves."""
import numpy as np


def f16(x):
    return np.asarray(x, dtype=np.float64).astype(np.float16)


def bits(h):
    return np.asarray(h, dtype=np.float16).view(np.uint16)


def sat(x):
    return np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)


def test_message_with_exponent_0x6400_is_1024_plus_c_and_hadd2_gives_v():
    c = np.arange(0, 255, dtype=np.uint16)                     # C = 127 - v, v in [-127, 127]
    as_half = (c + np.uint16(0x6400)).view(np.float16)
    assert np.array_equal(as_half.astype(np.float64), 1024.0 + c)
    v = f16(1151.0 - as_half.astype(np.float64))
    assert np.array_equal(v.astype(np.int64), 127 - c.astype(np.int64))
    assert bits(v[c == 127])[0] == 0                           # v = 0 is +0, never -0
    assert bits(np.float16(1151.0)) == 0x647F


def test_self_correction_rule_as_two_fmas():
    v, old = np.meshgrid(np.arange(-127, 128), np.arange(-127, 128), indexing="ij")
    keep = sat(f16(v.astype(np.float64) * old + 1.0).astype(np.float64))       # HFMA2.SAT(v, v_old, 1.0)
    assert set(np.unique(keep)) == {0.0, 1.0}
    cor = f16(v * keep + 0.0)                                                   # HFMA2(v, keep, +0)
    want = np.where(((v < 0) != (old < 0)) & (old != 0), 0, v)                 # :422-426 (a zero counts as non-negative)
    assert np.array_equal(cor.astype(np.int64), want)
    assert not (bits(cor)[want == 0] & 0x8000).any()                           # killed or zero -> +0: sign bit clear


def test_sign_of_u_by_one_fma_with_the_1536_constant():
    mu = np.arange(0, 128)
    for sign in (1.0, -1.0):
        x = f16(mu * sign + 1536.0)                                             # HFMA2(mu, +-1.0, 1536)
        assert np.array_equal(x.astype(np.float64), 1536.0 + sign * mu)         # ulp is 1 in [1024, 2048)
        u = (bits(x).astype(np.uint32) + 0x9A00) & 0xFFFF                       # VIADD.16x2 with -0x6600
        assert np.array_equal(u.astype(np.uint16).view(np.int16).astype(np.int64), (sign * mu).astype(np.int64))
    assert bits(np.float16(1536.0)) == 0x6600 and bits(np.float16(1.0)) == 0x3C00


def test_sign_word_is_plus_or_minus_one():
    # (XOR of the row's lanes & 0x8000) ^ 0x3c00, then ^ (own lane & 0x8000): bit 15 = parity of the other signs
    for total in (0, 1):
        for own in (0, 1):
            w = ((total << 15) ^ 0x3C00) ^ (own << 15)
            assert np.uint16(w).view(np.float16) == (-1.0 if total ^ own else 1.0)
