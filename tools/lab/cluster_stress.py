"""Repeated parity runs of the cluster kernel variants against the oracle (development aid)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, os, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests")); sys.path.insert(0, os.path.join(%r, "oracle"))
import labrador_ldpc_b200 as L
from pyoracle import Oracle
import test_gpu_k16384 as T
o = Oracle()
bad = 0
for code in (9, 10, 11):
    for seed in range(int(sys.argv[1])):
        _, llrs = T.frames(L, code, 90, T.EBN0[code], seed=1000 + 17 * seed + code, ty="i8")
        want = o.decode_ms_batch(code, llrs, 40, nthreads=16)
        for rep in range(2):
            got = L.LDPCCode(code).decode_ms_batch(llrs, 40)
            m = [int((np.asarray(g) != np.asarray(w)).any(axis=tuple(range(1, np.asarray(g).ndim))).sum()) if np.asarray(g).ndim > 1 else int((np.asarray(g).astype(np.int64) != np.asarray(w).astype(np.int64)).sum()) for g, w in zip(got, want)]
            if any(m):
                bad += 1
                print("MISMATCH code", code, "seed", seed, "rep", rep, m, flush=True)
print("runs with mismatches:", bad)
''' % (ROOT, ROOT, ROOT)
for env in ({"LABRADOR_LDPC_CLUSTER_ASYNC": "0"}, {}, {"LABRADOR_LDPC_CLUSTER_ASYNC": "0", "LABRADOR_LDPC_CLUSTER_MINB": "1"}):
    print("== env", env, flush=True)
    subprocess.call([sys.executable, "-c", CHILD, sys.argv[1] if len(sys.argv) > 1 else "4"],
                    env=dict(os.environ, LABRADOR_LDPC_NO_REBUILD="1", **env))
