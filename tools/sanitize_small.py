"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import labrador_ldpc_b200 as L
import pyoracle
from frames import make_frames, hard_frames
o = pyoracle.Oracle()
for code, ty, eb in ((8, "i8", 2.0), (5, "i8", 2.0), (6, "i8", 4.0), (5, "f32", 2.0), (3, "i8", 4.0), (0, "i8", 3.0), (2, "f32", 3.0)):
    c = L.LDPCCode(code)
    _, _, llrs = make_frames(o, code, 24, eb, seed=1, ty=ty)
    want = o.decode_ms_batch(code, llrs, 30, nthreads=4)
    got = c.decode_ms_batch(llrs, 30)
    assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), (code, ty)
for code in (0, 5, 8):
    c = L.LDPCCode(code)
    d, cw, rx = hard_frames(o, code, 16, 3, seed=2)
    got = c.decode_bf_batch(rx, 20); want = o.decode_bf_batch(code, rx, 20)
    assert all(np.array_equal(np.asarray(g).astype(np.int64), np.asarray(w).astype(np.int64)) for g, w in zip(got, want)), code
    assert np.array_equal(c.copy_encode_batch(d), cw)
    l = c.hard_to_llrs_batch(cw, "f32"); assert np.array_equal(c.llrs_to_hard_batch(l), cw)
print("sanitize_small OK")
