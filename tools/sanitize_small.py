"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import labrador_ldpc_b200 as L
import pyoracle
from frames import make_frames, hard_frames, soft_frames, quantise_soft
o = pyoracle.Oracle()
for code, ty, eb in ((8, "i8", 2.0), (5, "i8", 2.0), (6, "i8", 4.0), (5, "f32", 2.0), (3, "i8", 4.0), (0, "i8", 3.0), (2, "f32", 3.0)):
    c = L.LDPCCode(code)
    _, _, llrs = make_frames(o, code, 24, eb, seed=1, ty=ty)
    want = o.decode_ms_batch(code, llrs, 30, nthreads=4)
    got = c.decode_ms_batch(llrs, 30)
    assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), (code, ty)
for code in (0, 1, 2, 3, 5, 6, 8):
    c = L.LDPCCode(code)
    d, cw, rx = hard_frames(o, code, 16, 3, seed=2)
    got = c.decode_bf_batch(rx, 20); want = o.decode_bf_batch(code, rx, 20)
    assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), code
    assert np.array_equal(c.copy_encode_batch(d), cw)
    l = c.hard_to_llrs_batch(cw, "f32"); assert np.array_equal(c.llrs_to_hard_batch(l), cw)
same = lambda got, want: all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want))
for code, eb in ((0, 3.0), (3, 3.6), (5, 2.0), (8, 2.0)):                       # fused front ends
    c = L.LDPCCode(code)
    _, soft = soft_frames(o, code, 16, eb, seed=3)
    assert same(c.decode_ms_soft_batch(soft, 4.0, 31, 30, "i8"), o.decode_ms_batch(code, quantise_soft(soft, 4.0, 31, np.int8), 30, nthreads=4))
    assert same(c.decode_ms_soft_batch(soft, 256.0, 8191, 30, "i16"), o.decode_ms_batch(code, quantise_soft(soft, 256.0, 8191, np.int16), 30, nthreads=4))
    _, _, rx = hard_frames(o, code, 9, 2, seed=4)
    assert same(c.decode_ms_hard_batch(rx, 30), o.decode_ms_batch(code, np.stack([o.hard_to_llrs(code, r, "i8") for r in rx]), 30, nthreads=4))
for code in (0, 5):                                                              # harness kernels
    c = L.LDPCCode(code)
    data = c.random_data_batch(5, 0, np.zeros((13, c.k() // 8), np.uint8))
    cw = c.copy_encode_batch(data)
    for ty, lim in (("f32", 0), ("i8", 31), ("i16", 8191)):
        c.awgn_batch(cw, 0.7, 9.0, 5, 0, ty, limit=lim)
    out = np.zeros((13, c.output_len()), np.uint8); out[:, : c.n() // 8] = cw
    assert not c.count_errors_batch(out, data).any()
for code in range(9):                                                            # encoders, ragged batch, in place too
    c = L.LDPCCode(code)                                                         # (LABRADOR_LDPC_ENC_TM_FORM=1: compact TM form; LABRADOR_LDPC_ENC_TC_COPIES=1: table copies for TC512 too; LABRADOR_LDPC_ENC_GENERATOR=1: generator kernel)
    d = np.random.default_rng(code).integers(0, 256, (37, c.k() // 8), dtype=np.uint8)
    want = o.copy_encode_batch(code, d, nthreads=4)
    assert np.array_equal(c.copy_encode_batch(d), want)
    ip = np.zeros_like(want); ip[:, : c.k() // 8] = d
    c.encode_batch(ip)
    assert np.array_equal(ip, want)
# round 2: i16 on the packed-check-side kernel (all five TM codes it serves), i32 / f64 on the scalar-lane kernel with
# the in-thread exit test, TM1280 i8 on the packed kernel (covered above), the k = 16384 codes (encode, bf, min-sum)
for code, ty, eb in ((4, "i16", 2.6), (5, "i16", 1.8), (6, "i16", 3.4), (7, "i16", 2.4), (8, "i16", 1.6), (5, "i32", 1.8), (6, "f64", 3.4), (3, "i16", 3.6)):
    c = L.LDPCCode(code)
    _, _, llrs = make_frames(o, code, 20, eb, seed=7, ty=ty)
    assert same(c.decode_ms_batch(llrs, 30), o.decode_ms_batch(code, llrs, 30, nthreads=4)), (code, ty)
for code in (9, 10, 11):
    c = L.LDPCCode(code)
    d = np.random.default_rng(code).integers(0, 256, (5, c.k() // 8), dtype=np.uint8)
    cw = c.copy_encode_batch(d)
    rx = cw.copy(); rx[:, 3] ^= 0x41
    assert same(c.decode_bf_batch(rx, 10), o.decode_bf_batch(code, rx, 10))
    llrs = c.hard_to_llrs_batch(rx, "i8")
    assert same(c.decode_ms_batch(llrs, 6), o.decode_ms_batch(code, llrs, 6, nthreads=4))
print("sanitize_small OK")
