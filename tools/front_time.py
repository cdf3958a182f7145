"""Fused front end vs the two-step sequence (quantise kernel + decode_ms), device-resident f32 soft values.
usage: front_time.py [code] [batch] [ebn0]   -- development aid / profiles/r01_front.md"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import labrador_ldpc_b200 as L
from quick_time import gen

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if __name__ == "__main__":
    code = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    ebn0 = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
    c = L.LDPCCode(code)
    _, soft = gen(code, batch, ebn0, "f32")
    q = c.quantise_batch(soft, 4.0, 31, "i8")
    out, ok, it = c.decode_ms_batch(q, 100)
    out2, ok2, it2 = c.decode_ms_soft_batch(soft, 4.0, 31, 100, "i8")
    assert torch.equal(out, out2) and torch.equal(ok, ok2) and torch.equal(it, it2)
    t_q = timed(lambda: c.quantise_batch(soft, 4.0, 31, "i8", llrs=q))
    t_d = timed(lambda: c.decode_ms_batch(q, 100, output=out, success=ok, iters=it))
    t_f = timed(lambda: c.decode_ms_soft_batch(soft, 4.0, 31, 100, "i8", output=out, success=ok, iters=it))
    gb = batch * c.n() * 5 / 1e6
    print("%s batch %d: quantise %.3f ms (%.0f GB/s) + decode %.3f ms = %.3f ms two-step; fused %.3f ms (%.2f M cw/s, %.3fx)" % (
        c.name, batch, t_q, gb / t_q, t_d, t_q + t_d, t_f, batch / t_f / 1e3, (t_q + t_d) / t_f))
    cw = c.copy_encode_batch(torch.randint(0, 256, (batch, c.k() // 8), dtype=torch.uint8, device="cuda"))
    l8 = c.hard_to_llrs_batch(cw, "i8")
    t_h = timed(lambda: c.hard_to_llrs_batch(cw, "i8", llrs=l8))
    t_d = timed(lambda: c.decode_ms_batch(l8, 100, output=out, success=ok, iters=it))
    t_f = timed(lambda: c.decode_ms_hard_batch(cw, 100, output=out, success=ok, iters=it))
    print("%s batch %d clean hard input: hard_to_llrs %.3f ms + decode %.3f ms = %.3f ms two-step; fused %.3f ms (%.3fx)" % (
        c.name, batch, t_h, t_d, t_h + t_d, t_f, (t_h + t_d) / t_f))
