import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import labrador_ldpc_b200 as L
from pyoracle import Oracle
import test_gpu_k16384 as T
o = Oracle()
code = 9
_, llrs = T.frames(L, code, 6, T.EBN0[code], seed=1009, ty="i8")
for mi in (1, 2, 3, 5, 40):
    want = o.decode_ms_batch(code, llrs, mi, nthreads=8)
    got = L.LDPCCode(code).decode_ms_batch(llrs, mi)
    diff = [int(np.unpackbits(np.asarray(g) ^ np.asarray(w)).sum()) for g, w in zip(got[0], want[0])]
    print("maxiters", mi, "bit diffs per frame", diff, "ok", list(np.asarray(got[1]).astype(int)), list(np.asarray(want[1]).astype(int)),
          "iters", list(np.asarray(got[2])), list(np.asarray(want[2])), flush=True)
