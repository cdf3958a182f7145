"""Device-timed throughput of every operation x code (x LLR type) of the hot path.
Writes a markdown table (profiles/).  Development/measurement aid; bench.py is the contract harness."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import labrador_ldpc_b200 as L

EBN0 = {0: 3.0, 1: 3.0, 2: 2.5, 3: 4.0, 4: 3.0, 5: 2.0, 6: 4.0, 7: 3.0, 8: 2.0, 9: 3.6, 10: 2.6, 11: 1.8}


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def gen(c, batch, ebn0, seed=1):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    data = torch.randint(0, 256, (batch, c.k() // 8), dtype=torch.uint8, device="cuda", generator=g)
    cw = c.copy_encode_batch(data)
    bits = ((cw.unsqueeze(-1) >> torch.arange(7, -1, -1, device="cuda", dtype=torch.uint8)) & 1).reshape(batch, -1)
    sigma2 = 1.0 / (2.0 * (c.k() / c.n()) * 10.0 ** (ebn0 / 10.0))
    y = (1.0 - 2.0 * bits.float()) + (sigma2 ** 0.5) * torch.randn(bits.shape, device="cuda", generator=g)
    return data, cw, (2.0 / sigma2) * y


def main():
    out = ["# r02 -- throughput sweep on one B200 (`python tools/sweep.py`)", "",
           "Device-resident buffers, CUDA events, 3 repetitions after one warm-up; decode at the listed Eb/N0, max_iters 100;",
           "encode timed on min(2^24, 2^33 / n) codewords (1 GiB of codewords for the TM codes);",
           "bf input = codeword with 3 random bit flips.  Mcw/s = 1e6 codewords per second; Gbit/s counts information bits.", "",
           "| code | Eb/N0 | enc Mcw/s (GB/s out) | bf Mcw/s | ms i8 Mcw/s (Gbit/s, kernel) | ms i16 | ms i32 | ms f32 | ms f64 | h2l f32 GB/s | l2h f32 GB/s |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for code in range(12):
        c = L.LDPCCode(code)
        n, k = c.n(), c.k()
        batch = max(4096, min(1 << 18, (1 << 28) // n))
        if code >= 9: batch = 4096          # k = 16384: min-sum runs on the table-driven kernel (messages in an L2 slot)
        data, cw, llr = gen(c, batch, EBN0[code])
        # the encoders are fast enough that 2^18 codewords are launch-bound: time them on 1 GiB of codewords
        ebatch = max(batch, min(1 << 24, (1 << 33) // n))
        edata = torch.randint(0, 256, (ebatch, k // 8), dtype=torch.uint8, device="cuda")
        cw2 = torch.empty((ebatch, n // 8), dtype=torch.uint8, device="cuda")
        t_enc = timeit(lambda: c.copy_encode_batch(edata, cw2)) * batch / ebatch
        del edata, cw2
        rx = cw.clone()
        idx = torch.randint(0, n, (batch, 3), device="cuda")
        for j in range(3):
            rx[torch.arange(batch, device="cuda"), idx[:, j] // 8] ^= (1 << (7 - (idx[:, j] % 8))).to(torch.uint8)
        o = torch.empty((batch, c.output_len()), dtype=torch.uint8, device="cuda")
        ok = torch.empty(batch, dtype=torch.uint8, device="cuda"); it = torch.empty(batch, dtype=torch.int32, device="cuda")
        t_bf = timeit(lambda: c.decode_bf_batch(rx, 50, output=o, success=ok, iters=it))
        cells = []
        for ty in ("i8", "i16", "i32", "f32", "f64"):
            nb = batch if ty == "i8" else max(1024, batch // (8 if ty != "f64" else 32))
            if code >= 9: nb = 1024
            if ty == "i8": q = torch.clamp(torch.round(4 * llr[:nb]), -31, 31).to(torch.int8)
            elif ty == "i16": q = torch.clamp(torch.round(256 * llr[:nb]), -8191, 8191).to(torch.int16)
            elif ty == "i32": q = torch.round(65536 * llr[:nb]).to(torch.int32)
            elif ty == "f32": q = llr[:nb].contiguous()
            else: q = llr[:nb].double().contiguous()
            q = q.contiguous()
            t = timeit(lambda: c.decode_ms_batch(q, 100, output=o[:nb], success=ok[:nb], iters=it[:nb]))
            if ty == "i8":
                cells.append("%.2f (%.2f, %s)" % (nb / t / 1e6, nb * k / t / 1e9, c.decode_ms_kernel_name("i8")))
            else:
                cells.append("%.3f" % (nb / t / 1e6))
            del q
        hard = cw
        l32 = torch.empty((batch, n), dtype=torch.float32, device="cuda")
        t_h2l = timeit(lambda: c.hard_to_llrs_batch(hard, "f32", llrs=l32))
        h2 = torch.empty_like(hard)
        t_l2h = timeit(lambda: c.llrs_to_hard_batch(l32, output=h2))
        byt = batch * (n // 8 + 4 * n)
        out.append("| %s | %.1f | %.2f (%.1f) | %.2f | %s | %s | %s | %s | %s | %.0f | %.0f |" % (
            c.name, EBN0[code], batch / t_enc / 1e6, batch * (n // 8) / t_enc / 1e9, batch / t_bf / 1e6,
            cells[0], cells[1], cells[2], cells[3], cells[4], byt / t_h2l / 1e9, byt / t_l2h / 1e9))
        print(out[-1], flush=True)
        del data, cw, llr, rx, o, l32, h2
        torch.cuda.empty_cache()
    path = os.path.join(ROOT, "gpurun_out", "r02_sweep.md")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
