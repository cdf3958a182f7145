"""Device-side timing of decode_bf for every code (development aid).  usage: bf_time.py [batch] [flips]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import labrador_ldpc_b200 as L

if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    flips = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    codes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else range(9)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for code in codes:
        c = L.LDPCCode(code)
        b = batch if code < 6 else batch // 4
        data = torch.randint(0, 256, (b, c.k() // 8), dtype=torch.uint8, device="cuda", generator=g)
        cw = c.copy_encode_batch(data)
        for _ in range(flips):
            pos = torch.randint(0, c.n(), (b,), device="cuda", generator=g)
            cw[torch.arange(b, device="cuda"), pos // 8] ^= (128 >> (pos % 8)).to(torch.uint8)
        out, ok, it = c.decode_bf_batch(cw, 50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): c.decode_bf_batch(cw, 50, output=out, success=ok, iters=it)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        byt = b * (c.n() // 8 + c.output_len() + 5)
        print("%s bf batch %d flips %d: %.3f ms  %.1f M cw/s  %.0f GB/s  succ %.4f iters %.2f" % (
            c.name, b, flips, ms, b / ms / 1e3, byt / ms / 1e6, ok.float().mean().item(), it.float().mean().item()), flush=True)
