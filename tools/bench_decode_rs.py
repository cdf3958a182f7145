"""The reference's benches/decode.rs and benches/encode.rs as a device-timed harness.

benches/decode.rs:11-71 decodes ONE fixed vector per code -- data 0, 1, 2, ..., encoded, `rxcode[0] ^= 0xA8` (three bit
errors), 50 iterations at most -- with decode_bf, decode_ms::<i8> and decode_ms::<f32> (hard_to_llrs of the received word),
and reports ns per call; benches/encode.rs:11-23 reports copy_encode in MB/s of input (`b.bytes = k/8`).  Here the same
vector is replicated B times in device memory and decoded by one `_batch` launch (CUDA events, 5 repetitions); the CPU
column is the oracle (C++ restatement of the reference) on ONE core, the way `cargo bench` runs.  Writes a markdown table.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import labrador_ldpc_b200 as L
import pyoracle


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def cpu_ns(fn, per_call, min_s=0.3):
    """ns per codeword of `fn`, which processes `per_call` copies of the vector on ONE thread (the ctypes / numpy
    overhead of a call is amortised over the copies, as a Rust bench loop has none)."""
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < min_s:
        fn(); n += 1
    return (time.perf_counter() - t0) / (n * per_call) * 1e9


def main():
    o = pyoracle.Oracle(native=True)
    out = ["# r02 -- the reference's benches (benches/decode.rs, benches/encode.rs) on one B200 and on one host core", "",
           "Fixed vector of benches/decode.rs:17-21 (data 0,1,2,..., `rxcode[0] ^= 0xA8`, max 50 iterations), replicated B times and",
           "decoded by one `_batch` launch; ns per codeword = launch time / B.  CPU = the oracle (C++ restatement of the",
           "reference, -O3 -march=native) on one core.  Encode as benches/encode.rs: MB/s of input data (k/8 bytes per call).", "",
           "| code | B | bf GPU ns | bf CPU ns | ms i8 GPU ns | ms i8 CPU ns | ms f32 GPU ns | ms f32 CPU ns | iters (bf / i8 / f32) | encode GPU MB/s | encode CPU MB/s |",
           "|---|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|"]
    for code in range(9):
        c = L.LDPCCode(code)
        n, k = c.n(), c.k()
        data = (np.arange(k // 8) % 256).astype(np.uint8)
        cw = o.copy_encode(code, data)
        rx = cw.copy(); rx[0] ^= 0xA8
        B = max(8192, min(1 << 20, (1 << 29) // n))
        d_rx = torch.from_numpy(rx).cuda().repeat(B, 1).contiguous()
        res = {}
        obf, okbf, itbf = c.decode_bf_batch(d_rx, 50)
        assert bool(okbf.all()) and torch.equal(obf[0, : n // 8].cpu(), torch.from_numpy(cw))
        res["bf"] = timeit(lambda: c.decode_bf_batch(d_rx, 50, output=obf, success=okbf, iters=itbf)) / B * 1e9
        wok, wit, _ = o.decode_bf(code, rx, 50); assert wok and wit == int(itbf[0])
        R = max(4, 65536 // n)                       # copies per CPU call
        rxR = np.tile(rx, (R, 1))
        res["bf_cpu"] = cpu_ns(lambda: o.decode_bf_batch(code, rxR, 50, nthreads=1), R)
        iters = [int(itbf[0])]
        for ty in ("i8", "f32"):
            llr = c.hard_to_llrs_batch(d_rx, ty)
            om, okm, itm = c.decode_ms_batch(llr, 50)
            assert bool(okm.all()) and torch.equal(om[0, : n // 8].cpu(), torch.from_numpy(cw))
            res[ty] = timeit(lambda: c.decode_ms_batch(llr, 50, output=om, success=okm, iters=itm)) / B * 1e9
            l1 = llr[0].cpu().numpy()
            wok, wit, _ = o.decode_ms(code, l1, 50); assert wok and wit == int(itm[0])
            lR = np.tile(l1, (R, 1))
            res[ty + "_cpu"] = cpu_ns(lambda: o.decode_ms_batch(code, lR, 50, nthreads=1), R)
            iters.append(int(itm[0]))
            del llr, om
        EB = max(B, min(1 << 24, (1 << 31) // n))
        d_data = torch.from_numpy(data).cuda().repeat(EB, 1).contiguous()
        d_cw = torch.empty((EB, n // 8), dtype=torch.uint8, device="cuda")
        t_enc = timeit(lambda: c.copy_encode_batch(d_data, d_cw))
        assert torch.equal(d_cw[EB - 1].cpu(), torch.from_numpy(cw))
        dR = np.tile(data, (R, 1))
        enc_cpu = cpu_ns(lambda: o.copy_encode_batch(code, dR, nthreads=1), R)
        out.append("| %s | %d | %.2f | %.0f | %.2f | %.0f | %.2f | %.0f | %d / %d / %d | %.0f | %.1f |" % (
            c.name, B, res["bf"], res["bf_cpu"], res["i8"], res["i8_cpu"], res["f32"], res["f32_cpu"], iters[0], iters[1], iters[2],
            EB * (k // 8) / t_enc / 1e6, (k // 8) / enc_cpu * 1e3))
        print(out[-1], flush=True)
        del d_rx, d_data, d_cw
        torch.cuda.empty_cache()
    path = os.path.join(ROOT, "gpurun_out", "r02_reference_benches.md")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
