"""BASELINE.json configs[3]: TM5120 (4 dB) + TM6144 (3 dB) i8 mixed-code batch, half the frames each.
usage: mixed_time.py [frames_total]   -- development / profiles aid"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import labrador_ldpc_b200 as L

def frames(c, batch, ebn0, seed):
    sigma2 = 1.0 / (2.0 * (c.k() / c.n()) * 10.0 ** (ebn0 / 10.0))
    data = c.random_data_batch(seed, 0, torch.empty((batch, c.k() // 8), dtype=torch.uint8, device="cuda"))
    cw = c.copy_encode_batch(data)
    return c.awgn_batch(cw, sigma2 ** 0.5, 8.0 / sigma2, seed, 0, "i8", limit=31)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

if __name__ == "__main__":
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    a, b = L.LDPCCode.TM5120, L.LDPCCode.TM6144
    la, lb = frames(a, total // 2, 4.0, 11), frames(b, total // 2, 3.0, 12)
    ra, rb = a.decode_ms_batch(la, 100), b.decode_ms_batch(lb, 100)
    t_a = timed(lambda: a.decode_ms_batch(la, 100, output=ra[0], success=ra[1], iters=ra[2]))
    t_b = timed(lambda: b.decode_ms_batch(lb, 100, output=rb[0], success=rb[1], iters=rb[2]))
    t_m = timed(lambda: L.decode_ms_mixed([(a, la), (b, lb)], 100))
    m = L.decode_ms_mixed([(a, la), (b, lb)], 100)
    torch.cuda.synchronize()
    assert torch.equal(m[0][0], ra[0]) and torch.equal(m[1][0], rb[0]) and torch.equal(m[0][2], ra[2]) and torch.equal(m[1][2], rb[2])
    gbit = lambda ms: total * 4096 / ms / 1e6
    print("TM5120 @4 dB: %d frames %.3f ms (%.2f M cw/s, FER %.1e, iters %.2f)" % (total // 2, t_a, total / 2 / t_a / 1e3, 1 - ra[1].float().mean().item(), ra[2].float().mean().item()))
    print("TM6144 @3 dB: %d frames %.3f ms (%.2f M cw/s, FER %.1e, iters %.2f)" % (total // 2, t_b, total / 2 / t_b / 1e3, 1 - rb[1].float().mean().item(), rb[2].float().mean().item()))
    print("back to back: %.3f ms = %.2f Gbit/s;  mixed (one stream per code): %.3f ms = %.2f Gbit/s, %.2f M cw/s" % (
        t_a + t_b, gbit(t_a + t_b), t_m, gbit(t_m), total / t_m / 1e3))
