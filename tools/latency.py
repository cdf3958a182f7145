"""Latency of the reference-signature single-codeword calls (host buffers, blocking) -- development aid."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import labrador_ldpc_b200 as L

for code in (0, 2, 5, 8):
    c = L.LDPCCode(code)
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, c.k() // 8, dtype=np.uint8)
    cw = np.zeros(c.n() // 8, np.uint8)
    c.copy_encode(data, cw)
    llrs = np.zeros(c.n(), np.int8)
    c.hard_to_llrs(cw, llrs)
    llrs = (llrs * 8).astype(np.int8)
    llrs[::7] = -llrs[::7] // 4                         # a few weak wrong bits: needs several iterations
    out = np.zeros(c.output_len(), np.uint8)
    res = {}
    for name, fn in (("copy_encode", lambda: c.copy_encode(data, cw)),
                     ("decode_ms_i8", lambda: c.decode_ms(llrs, out, maxiters=50)),
                     ("decode_bf", lambda: c.decode_bf(cw, out, maxiters=50))):
        for _ in range(20): fn()
        t = time.perf_counter()
        reps = 300
        for _ in range(reps): fn()
        res[name] = (time.perf_counter() - t) / reps * 1e6
    ok, it = c.decode_ms(llrs, out, maxiters=50)
    print("%s single-codeword call latency (us): %s   [decode_ms: ok=%s iters=%d]" % (
        c.name, ", ".join("%s %.1f" % kv for kv in res.items()), ok, it), flush=True)
