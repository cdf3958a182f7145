#!/usr/bin/env python3
"""Device-timed harness for the batched LDPC hot path (BASELINE.json metric and configs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

Workloads = BASELINE.json `configs` (SURVEY.md section 8d gives their shapes):

  c1  TC128 decode_ms, i8 LLRs, 100 000 frames at each of Eb/N0 = 0..4 dB (the reference's CPU-runnable case)
  c2  TM2048 decode_ms i16 + decode_ms f32 + decode_bf, 1 Mi frames each, punctured column exercised
  c3  TM8192 decode_ms i8 at 2 dB, 2 Mi frames per GPU (= the 16 Mi-frame config on 8 GPUs)       [default, headline]
  c4  TM5120 at 4 dB + TM6144 at 3 dB decode_ms i8, mixed-code batch, half the frames each, one stream per code
  c5  batched copy_encode of all nine codes (reference benches/encode.rs, throughput of input bytes)

Weak scaling: every GPU processes the same-sized shard of independent codewords; no collective on the data path.
A "step" is one pass of every part of the workload over the rank's shard, inputs resident in HBM.

`value` is the whole-job throughput in the workload's unit (decode: information Gbit/s over all frames, FER beside
it; encode: MB/s of input data as the reference's `b.bytes = k/8`).  `e2e` is the same metric through the C ABI
with HOST (pinned) buffers, copies inside the timed region, plus a copy-only control of the same transport.
`cpu_baseline` is the CPU oracle (C++ restatement of the reference; the Rust crate cannot be built in this image) on
the box's host cores over a bounded prefix of the very same input bytes -- which doubles as the exact-parity check.

`--impl reference` times the reference's CPU algorithm (the oracle) alone on the same workload.
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_ITERS = 100
NAMES = ["TC128", "TC256", "TC512", "TM1280", "TM1536", "TM2048", "TM5120", "TM6144", "TM8192"]
LLR_BYTES = {"i8": 1, "i16": 2, "i32": 4, "f32": 4, "f64": 8}
QUANT = {"i8": (4.0, 31), "i16": (256.0, 8191), "f32": (1.0, 0)}      # scale, clamp
# tracked ncu export of the shipped headline kernel (tools/ncu_export.py writes both)
NCU_CSV = os.path.join(ROOT, "profiles", "r02_tm8192_ncu.csv")
NCU_META = os.path.join(ROOT, "profiles", "r02_tm8192_ncu.json")


def workload_parts(name, scale=1.0):
    """The parts of one step.  `scale` multiplies every frame count (--frames-per-gpu)."""
    f = lambda x: max(1, int(x * scale))
    if name == "c1":
        return [dict(op="ms", code=0, ty="i8", ebn0=float(e), frames=f(100_000)) for e in range(5)]
    if name == "c2":
        return [dict(op="ms", code=5, ty="i16", ebn0=2.0, frames=f(1 << 20)),
                dict(op="ms", code=5, ty="f32", ebn0=2.0, frames=f(1 << 20)),
                dict(op="bf", code=5, ty="i8", ebn0=9.0, frames=f(1 << 20))]
    if name == "c3":
        return [dict(op="ms", code=8, ty="i8", ebn0=2.0, frames=f(1 << 21))]
    if name == "c4":
        return [dict(op="ms", code=6, ty="i8", ebn0=4.0, frames=f(1 << 20)),
                dict(op="ms", code=7, ty="i8", ebn0=3.0, frames=f(1 << 20))]
    if name == "c5":
        n_of = [128, 256, 512, 1280, 1536, 2048, 5120, 6144, 8192]
        return [dict(op="enc", code=c, ty="u8", ebn0=None, frames=f((256 << 20) // (n_of[c] // 8))) for c in range(9)]
    raise SystemExit("unknown workload " + name)


WORKLOAD_TEXT = {
    "c1": "TC128 (n=128, r=1/2) decode_ms i8 LLRs, 100k frames at each of Eb/N0 0..4 dB, max_iters 100",
    "c2": "TM2048 (k=1024, r=1/2) decode_ms i16 + f32 at 2 dB and decode_bf, 1 Mi frames each, max_iters 100",
    "c3": "TM8192 (k=4096, r=1/2) decode_ms i8 LLRs, Eb/N0 2 dB, max_iters 100",
    "c4": "TM5120 (r=4/5, 4 dB) + TM6144 (r=2/3, 3 dB) decode_ms i8 mixed-code batch, half the frames each, max_iters 100",
    "c5": "copy_encode sweep over all nine codes, bit-packed, 256 MiB of codewords per code",
}
CONCURRENT = {"c4"}             # parts run on one stream each (labrador_ldpc_b200.decode_ms_mixed)
SMALL_INPUT = {"c1"}            # inputs smaller than L2: flush L2 between the timed steps


def part_name(p):
    tag = {"ms": "decode_ms<%s>" % p["ty"], "bf": "decode_bf", "enc": "copy_encode"}[p["op"]]
    return "%s %s%s" % (NAMES[p["code"]], tag, "" if p["ebn0"] is None else " @%.0f dB" % p["ebn0"])


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def read_ncu():
    """Metrics of the tracked `ncu --set full` capture of the shipped TM8192 i8 kernel (profiles/r02_tm8192_ncu.csv,
    raw export) and the launch it captured (profiles/r02_tm8192_ncu.json).  None if not present."""
    try:
        with open(NCU_CSV) as f:
            rows = list(csv.reader(f))
        with open(NCU_META) as f:
            meta = json.load(f)
        col = {name: i for i, name in enumerate(rows[0])}
        val = lambda k: float(rows[2][col[k]].replace(",", ""))
        unit = lambda k: rows[1][col[k]]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum(val(k) * mult[unit(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        return {"dram_bytes": dram, "frames": meta["frames"], "edge_updates": meta["edge_updates"],
                "alu_busy": val("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
                "sm_cycles": val("sm__cycles_active.avg"), "kernel": meta.get("kernel", ""),
                "issue_active": val("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
                "inst_executed": val("smsp__inst_executed.sum"),
                "fmaheavy_busy": val("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed") / 100.0}
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout_s=3.0):
        """nvidia-smi needs a moment to start: block until its first sample so that short timed regions are covered."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout_s:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        # samples inside the timed region; a region shorter than the sampling period takes the samples around it
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or \
               [r for (t, r) in self.rows if t0 - 0.25 <= t <= t1 + 0.25 and len(r) >= 7] or \
               [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


def sigma2_of(k, n, ebn0_db):
    return 1.0 / (2.0 * (k / n) * 10.0 ** (ebn0_db / 10.0))


def generate_part(L, torch, p, seed, first_frame):
    """Inputs of one part for this rank, produced on the device by the library's own kernels (csrc/channel.cu,
    encoders): random data -> encode -> BPSK + AWGN -> LLRs.  The generator is counter-based (Philox keyed
    by `seed`, counter = frame index), so rank r producing frames [r*frames, (r+1)*frames) yields its slice of one
    logical run whatever the world size.  Returns the device tensor the timed call reads."""
    c = L.LDPCCode(p["code"])
    frames, n, k = p["frames"], c.n(), c.k()
    if p["op"] == "enc":
        data = torch.empty((frames, k // 8), dtype=torch.uint8, device="cuda")
        chunk = max(1, (256 << 20) // (k // 8))
        for f0 in range(0, frames, chunk):
            c.random_data_batch(seed, first_frame + f0, data[f0:f0 + chunk])
        return data
    ty = "i8" if p["op"] == "bf" else p["ty"]
    tdt = {"i8": torch.int8, "i16": torch.int16, "f32": torch.float32}[ty]
    scale, limit = QUANT[ty]
    llrs = torch.empty((frames, n), dtype=tdt, device="cuda")
    s2 = sigma2_of(k, n, p["ebn0"])
    chunk = 65536
    data = torch.empty((chunk, k // 8), dtype=torch.uint8, device="cuda")
    cw = torch.empty((chunk, n // 8), dtype=torch.uint8, device="cuda")
    for f0 in range(0, frames, chunk):
        nf = min(chunk, frames - f0)
        c.random_data_batch(seed, first_frame + f0, data[:nf])
        c.copy_encode_batch(data[:nf], cw[:nf])
        c.awgn_batch(cw[:nf], s2 ** 0.5, scale * 2.0 / s2, seed, first_frame + f0, ty, limit=limit, out=llrs[f0:f0 + nf])
    if p["op"] == "bf":                               # bf input: llrs_to_hard of the channel output
        hard = c.llrs_to_hard_batch(llrs)
        torch.cuda.synchronize()
        return hard
    torch.cuda.synchronize()
    return llrs


class Part:
    """One homogeneous batch of a step: device-resident inputs, result buffers and the call that processes them."""

    def __init__(self, L, torch, p, seed, rank):
        self.p, self.L, self.torch = p, L, torch
        self.c = L.LDPCCode(p["code"])
        self.frames = p["frames"]
        self.inp = generate_part(L, torch, p, seed, rank * p["frames"])
        c = self.c
        if p["op"] == "enc":
            self.out = torch.empty((self.frames, c.n() // 8), dtype=torch.uint8, device="cuda")
            self.ok = self.iters = None
        else:
            self.out = torch.empty((self.frames, c.output_len()), dtype=torch.uint8, device="cuda")
            self.ok = torch.empty((self.frames,), dtype=torch.uint8, device="cuda")
            self.iters = torch.empty((self.frames,), dtype=torch.int32, device="cuda")

    # algorithmic HBM bytes per frame
    def alg_bytes_per_frame(self):
        c, p = self.c, self.p
        if p["op"] == "ms":
            return c.n() * LLR_BYTES[p["ty"]] + c.output_len() + 8
        if p["op"] == "bf":
            return c.n() // 8 + c.output_len() + 8
        return c.k() // 8 + c.n() // 8

    def info_bits(self):
        return self.frames * self.c.k()

    def run(self, inp=None, out=None, ok=None, iters=None, stream=None):
        """The timed call: device tensors (stream-ordered) or host arrays (blocking, through the pipeline)."""
        inp = self.inp if inp is None else inp
        out = self.out if out is None else out
        ok = self.ok if ok is None else ok
        iters = self.iters if iters is None else iters
        if self.p["op"] == "ms":
            self.c.decode_ms_batch(inp, MAX_ITERS, output=out, success=ok, iters=iters, stream=stream)
        elif self.p["op"] == "bf":
            self.c.decode_bf_batch(inp, MAX_ITERS, output=out, success=ok, iters=iters, stream=stream)
        else:
            self.c.copy_encode_batch(inp, out, stream=stream)

    def oracle_run(self, oracle, host_inp, nthreads):
        if self.p["op"] == "ms":
            return oracle.decode_ms_batch(self.p["code"], host_inp, MAX_ITERS, nthreads=nthreads)
        if self.p["op"] == "bf":
            return oracle.decode_bf_batch(self.p["code"], host_inp, MAX_ITERS, nthreads=nthreads)
        return (oracle.copy_encode_batch(self.p["code"], host_inp, nthreads=nthreads),)


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: run this rank (and allocate its pinned buffers) on the CPU cores NVML reports as local to
    its GPU, so the end-to-end leg's host<->device copies do not cross sockets.  Best effort; returns the core count."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    try:
        return pyoracle.Oracle(native=True), "native"
    except Exception:
        return pyoracle.Oracle(native=False), "generic"


def timed_oracle(oracle, part, host_inp, nthreads, budget_s):
    """Runs the oracle on a prefix of `host_inp` sized for about `budget_s` seconds.  Returns (frames, secs, outputs)."""
    n = len(host_inp)
    probe = min(n, max(4 * nthreads, 64))
    t = time.perf_counter()
    out = part.oracle_run(oracle, host_inp[:probe], nthreads)
    dt = max(time.perf_counter() - t, 1e-4)
    frames = int(min(n, max(probe, budget_s * probe / dt)))
    if frames == probe and dt > 0.5 * budget_s:
        return probe, dt, out
    t = time.perf_counter()
    out = part.oracle_run(oracle, host_inp[:frames], nthreads)
    return frames, time.perf_counter() - t, out


def unit_of(workload):
    return ("encode_input_mbyte_per_s", "MB/s") if workload == "c5" else ("decoded_info_gbit_per_s", "Gbit/s")


def to_value(workload, info_bits, seconds):
    return info_bits / 8.0 / seconds / 1e6 if workload == "c5" else info_bits / seconds / 1e9


def mix_rate(workload, parts_spec, rates_cw_s, k_of):
    """Throughput of the workload's own mix of parts given each part's codeword rate: total information over the time
    one step would take."""
    secs = sum(p["frames"] / r for p, r in zip(parts_spec, rates_cw_s))
    bits = sum(p["frames"] * k_of[p["code"]] for p in parts_spec)
    return to_value(workload, bits, secs)


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (oracle) alone, all host threads, same workload mix.  The inputs
    are numpy-generated frames (tests/frames.py); each step runs a bounded sample of every part."""
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    oracle, kind = load_oracle()
    from frames import make_frames
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    gen = pyoracle.Oracle(native=False)
    cores = len(os.sched_getaffinity(0))
    spec = workload_parts(args.workload)
    k_of = {c: gen.k(c) for c in range(9)}

    class RefPart:
        def __init__(self, p):
            self.p = p

        def inputs(self, frames):
            p = self.p
            rng = np.random.default_rng(7 + p["code"])
            if p["op"] == "enc":
                return rng.integers(0, 256, (frames, gen.k(p["code"]) // 8), dtype=np.uint8)
            ty = "i8" if p["op"] == "bf" else p["ty"]
            _, _, llrs = make_frames(gen, p["code"], frames, p["ebn0"], seed=1 + p["code"], ty=ty)
            if p["op"] == "bf":
                return np.packbits(llrs < 0, axis=1)
            return llrs

        def oracle_run(self, orc, x, nthreads):
            p = self.p
            if p["op"] == "ms":
                return orc.decode_ms_batch(p["code"], x, MAX_ITERS, nthreads=nthreads)
            if p["op"] == "bf":
                return orc.decode_bf_batch(p["code"], x, MAX_ITERS, nthreads=nthreads)
            return (orc.copy_encode_batch(p["code"], x, nthreads=nthreads),)

    parts = [RefPart(p) for p in spec]
    # size every part's sample so that one step is ~3 s of CPU work split evenly over the parts
    budget = 3.0 / len(parts)
    samples = []
    for rp in parts:
        probe = rp.inputs(max(4 * cores, 64))
        rp.oracle_run(oracle, probe, cores)                      # cold start (threads, page faults)
        t = time.perf_counter()
        rp.oracle_run(oracle, probe, cores)
        rate = len(probe) / max(time.perf_counter() - t, 1e-4)
        frames = int(max(len(probe), min(rate * budget, 1 << 18)))
        samples.append(rp.inputs(frames) if frames > len(probe) else probe)
    for _ in range(args.warmup):
        for rp, x in zip(parts, samples):
            rp.oracle_run(oracle, x, cores)
    secs = [0.0] * len(parts)
    last = [None] * len(parts)
    for _ in range(args.steps):
        for i, (rp, x) in enumerate(zip(parts, samples)):
            t = time.perf_counter()
            last[i] = rp.oracle_run(oracle, x, cores)
            secs[i] += time.perf_counter() - t
    rates = [len(x) * args.steps / s for x, s in zip(samples, secs)]
    value = mix_rate(args.workload, spec, rates, k_of)
    metric, unit = unit_of(args.workload)
    step_frames = sum(len(x) for x in samples)
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sum(secs) / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": spec[0]["ty"], "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "frames_per_step": step_frames,
                   "note": "C++ restatement of labrador-ldpc 1.2.1 (oracle/oracle.cpp, build: %s) on host cores; the Rust "
                           "toolchain is absent.  value = throughput of the workload's own mix of parts at the measured "
                           "per-part rates" % kind},
        "parts": [{"part": part_name(p), "sample_frames": len(x), "codewords_per_s": r}
                  for p, x, r in zip(spec, samples, rates)],
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": "%d frames per step over %d part(s), %d steps, one codec per host thread"
                                   % (step_frames, len(parts), args.steps)},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if args.workload == "c3":
        line["codewords_per_s"] = rates[0]
        line["fer"] = float(1.0 - last[0][1].mean())
        line["mean_iters"] = float(last[0][2].mean())
    print(json.dumps(line))


def split_leg(args):
    """Hidden mode (run by rank 0 in a fresh process while the other ranks wait): ONE process drives `--gpus` devices
    through the library's own multi-device path -- labrador_ldpc_cuda_init(devices) + one host-pointer batch, which the
    library splits into contiguous per-device shards (csrc/runtime.cu: run_host_batch_at) -- and checks the result
    against device 0 decoding the same frames alone."""
    import torch
    import labrador_ldpc_b200 as L
    n = args.gpus
    p = workload_parts(args.workload)[0]
    if p["op"] != "ms":
        print(json.dumps({"skipped": "first part is not a decode_ms part"}))
        return
    p = dict(p, frames=args.split_frames)
    torch.cuda.set_device(0)
    L.init([0])
    pt = Part(L, torch, p, seed=1, rank=0)
    pt.run()
    torch.cuda.synchronize()
    h_in = torch.empty(pt.inp.shape, dtype=pt.inp.dtype).pin_memory()
    h_in.copy_(pt.inp)
    h_out = torch.empty(pt.out.shape, dtype=torch.uint8).pin_memory()
    h_ok = torch.empty((pt.frames,), dtype=torch.uint8).pin_memory()
    h_it = torch.empty((pt.frames,), dtype=torch.int32).pin_memory()
    L.init(list(range(n)))
    launches0 = L.kernel_launch_count()
    pt.run(h_in, h_out, h_ok, h_it)                        # warm-up: contexts and tables of the other devices
    per_call = L.kernel_launch_count() - launches0
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        pt.run(h_in, h_out, h_ok, h_it)
    secs = (time.perf_counter() - t0) / reps
    match = bool(torch.equal(h_out, pt.out.cpu()) and torch.equal(h_it, pt.iters.cpu()) and torch.equal(h_ok, pt.ok.cpu()))
    print(json.dumps({"devices": n, "frames": pt.frames, "seconds": secs,
                      "value": to_value(args.workload, pt.info_bits(), secs), "unit": unit_of(args.workload)[1],
                      "kernel_launches_per_call": int(per_call), "matches_single_device_result": match,
                      "what": "one process, labrador_ldpc_cuda_init([0..%d]), one host-pointer decode_ms_batch call" % (n - 1)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--frames-per-gpu", type=int, default=0, help="frames of the first part (others scale with it)")
    ap.add_argument("--e2e-bytes", type=int, default=1 << 30, help="host input bytes of the end-to-end leg, per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split-leg", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--split-frames", type=int, default=1 << 16, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.split_leg:
        split_leg(args)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import labrador_ldpc_b200 as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    L.init([local_rank])

    wl = args.workload
    base = workload_parts(wl)
    scale = args.frames_per_gpu / base[0]["frames"] if args.frames_per_gpu else 1.0
    spec = workload_parts(wl, scale)
    parts = [Part(L, torch, p, seed=1 + i, rank=rank) for i, p in enumerate(spec)]
    concurrent = wl in CONCURRENT
    side = [torch.cuda.Stream() for _ in parts] if concurrent else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if wl in SMALL_INPUT else None

    def step(marks=None):
        cur = torch.cuda.current_stream()
        if concurrent:                                   # one stream per code, as labrador_ldpc_b200.decode_ms_mixed
            for pt, st in zip(parts, side):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    pt.run(stream=st.cuda_stream)
            for st in side:
                cur.wait_stream(st)
            return
        for i, pt in enumerate(parts):
            pt.run()
            if marks is not None:
                marks[i].record()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    sampler.wait_first()
    barrier()
    launches0 = L.kernel_launch_count()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    starts = [ev() for _ in range(args.steps)]
    marks = [[ev() for _ in parts] for _ in range(args.steps)]
    ends = [ev() for _ in range(args.steps)]
    t_wall0 = time.time()
    for s in range(args.steps):
        if flush is not None:
            flush.fill_(s & 0xFF)                        # inputs fit in L2: evict them between timed steps (not timed)
        starts[s].record()
        step(None if concurrent else marks[s])
        ends[s].record()
    barrier()
    t_wall1 = time.time()
    launches = L.kernel_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = sum(a.elapsed_time(b) for a, b in zip(starts, ends))
    part_ms = [0.0] * len(parts)
    if not concurrent:
        for s in range(args.steps):
            prev = starts[s]
            for i in range(len(parts)):
                part_ms[i] += prev.elapsed_time(marks[s][i])
                prev = marks[s][i]
    t = torch.tensor([ms_total] + part_ms, device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t[0].item())
    part_ms = [float(x) / args.steps for x in t[1:].tolist()]
    ms_per_step = ms_total / args.steps

    # per-part statistics, summed over the ranks
    st_rows = []
    for pt in parts:
        if pt.ok is None:
            st_rows.append(torch.zeros(3, device="cuda", dtype=torch.float64))
        else:
            it = pt.iters.double()
            st_rows.append(torch.stack([pt.ok.double().sum(), it.sum(),
                                        torch.where(pt.ok.bool(), it + 1.0, it).sum()]))
    stats = torch.stack(st_rows)
    if dist is not None:
        dist.all_reduce(stats)
    stats = stats.cpu().numpy()
    total_bits = sum(pt.info_bits() for pt in parts) * world
    value = to_value(wl, total_bits, ms_per_step * 1e-3)
    metric, unit = unit_of(wl)

    # ---- end-to-end through the C ABI with pinned HOST buffers (+ copy-only control of the same transport) ----
    in_bytes_per_step = sum(pt.inp[0].numel() * pt.inp.element_size() * pt.frames for pt in parts)
    e2e_scale = min(1.0, args.e2e_bytes / in_bytes_per_step)
    e2e = []
    for pt in parts:
        ef = max(1, int(pt.frames * e2e_scale))
        h = {"ef": ef, "inp": torch.empty((ef,) + tuple(pt.inp.shape[1:]), dtype=pt.inp.dtype).pin_memory(),
             "out": torch.empty((ef, pt.out.shape[1]), dtype=torch.uint8).pin_memory()}
        h["inp"].copy_(pt.inp[:ef])
        if pt.ok is not None:
            h["ok"] = torch.empty((ef,), dtype=torch.uint8).pin_memory()
            h["it"] = torch.empty((ef,), dtype=torch.int32).pin_memory()
        e2e.append(h)
    torch.cuda.synchronize()

    def e2e_step():
        for pt, h in zip(parts, e2e):                    # host pointers -> streamed through the GPU in chunks
            pt.run(h["inp"], h["out"], h.get("ok"), h.get("it"))

    def copy_control_step():
        for pt, h in zip(parts, e2e):
            if pt.p["op"] == "ms":
                pt.c.copy_control_batch(h["inp"], h["out"], h["ok"], h["it"])

    def wall(fn, reps):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        s = (time.perf_counter() - t0) / reps
        ts = torch.tensor([s], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        return float(ts.item())

    e2e_s = wall(e2e_step, 3)
    e2e_match = all(torch.equal(h["out"], pt.out[:h["ef"]].cpu()) and
                    (pt.iters is None or torch.equal(h["it"], pt.iters[:h["ef"]].cpu())) for pt, h in zip(parts, e2e))
    e2e_bits = sum(h["ef"] * pt.c.k() for pt, h in zip(parts, e2e)) * world
    h2d = sum(h["inp"].numel() * h["inp"].element_size() for h in e2e)
    d2h = sum(h["out"].numel() + (h["ef"] * 5 if "ok" in h else 0) for h in e2e)
    e2e_line = {"value": to_value(wl, e2e_bits, e2e_s), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "seconds_per_step": e2e_s, "frames_per_gpu": [h["ef"] for h in e2e],
                "matches_device_resident_run": bool(e2e_match),
                "api": "labrador_ldpc_*_batch with pinned host pointers (wall clock around the blocking calls)"}
    if any(pt.p["op"] == "ms" for pt in parts):
        cc_s = wall(copy_control_step, 3)
        cc_h2d = sum(h["inp"].numel() * h["inp"].element_size() for pt, h in zip(parts, e2e) if pt.p["op"] == "ms")
        cc_bits = sum(h["ef"] * pt.c.k() for pt, h in zip(parts, e2e) if pt.p["op"] == "ms") * world
        e2e_line["copy_control"] = {
            "seconds_per_step": cc_s, "h2d_gbs_per_gpu": cc_h2d / cc_s / 1e9,
            "value_if_transport_only": to_value(wl, cc_bits, cc_s), "unit": unit,
            "what": "labrador_ldpc_copy_control_batch: the same arrays through the same chunked pipeline, no kernel"}
    if numa is not None:
        e2e_line["host_cores_local_to_gpu"] = numa
    if world > 1 and wl == "c3":
        # the library's own multi-device split, driven by ONE process (rank 0) while the other ranks wait
        barrier()
        if rank == 0:
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--split-leg", "--gpus", str(world),
                                    "--workload", wl], env=env, capture_output=True, text=True, timeout=600)
                e2e_line["in_library_split"] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as exc:
                e2e_line["in_library_split"] = {"error": str(exc)[:200]}
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = read_peaks()
    f_max = (clocks.get("sm_max_mhz") or 1965.0) * 1e6
    part_lines = []
    for i, (pt, p) in enumerate(zip(parts, spec)):
        row = {"part": part_name(p), "frames_per_gpu": pt.frames, "alg_bytes_per_frame": pt.alg_bytes_per_frame()}
        if not concurrent:
            ms = part_ms[i]
            row.update({"ms_per_step": ms, "codewords_per_s": pt.frames * world / (ms * 1e-3),
                        "value": to_value(wl, pt.info_bits() * world, ms * 1e-3), "unit": unit,
                        "hbm_gbs_per_gpu": pt.alg_bytes_per_frame() * pt.frames / (ms * 1e-3) / 1e9,
                        "hbm_frac": pt.alg_bytes_per_frame() * pt.frames / (ms * 1e-3) / 1e9 / peak})
        if pt.ok is not None:
            tf = pt.frames * world
            row.update({"fer": 1.0 - stats[i][0] / tf, "mean_iters": stats[i][1] / tf})
            if p["op"] == "ms":
                row["kernel"] = pt.c.decode_ms_kernel_name(p["ty"])
        part_lines.append(row)

    # roofline of the dominant part (the one that takes the largest share of the step)
    dom = 0 if concurrent else max(range(len(parts)), key=lambda i: part_ms[i])
    dpt = parts[dom]
    dom_ms = ms_per_step if concurrent else part_ms[dom]
    if concurrent:
        alg_bytes = sum(pt.alg_bytes_per_frame() * pt.frames for pt in parts)
    else:
        alg_bytes = dpt.alg_bytes_per_frame() * dpt.frames
    hbm_achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel_of": part_name(spec[dom]),
                "share_of_step": dom_ms / ms_per_step}
    line = {
        "metric": metric, "value": value, "unit": unit,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "+".join(sorted({p["ty"] for p in spec})), "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[wl], "id": wl,
                   "frames_per_gpu": [pt.frames for pt in parts], "frames_total": sum(pt.frames for pt in parts) * world,
                   "input_bytes_per_gpu": in_bytes_per_step,
                   "l2": ("inputs (%.2f GiB per step) far exceed the 126 MB L2" % (in_bytes_per_step / 2 ** 30))
                   if flush is None else "inputs fit in L2: a 256 MiB buffer is rewritten between the timed steps (not timed)",
                   "sharding": "independent codewords per rank, no collective",
                   "generator": "philox4x32-10 seeds 1.., frames rank*frames_per_gpu.. (csrc/channel.cu)"},
        "parts": part_lines,
        "e2e": e2e_line,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if wl == "c3":
        tf = parts[0].frames * world
        edge_updates = 2.0 * 30720 * float(stats[0][2])      # 2E(iters+1) on success, 2E*max_iters on failure
        upd_per_gpu_s = edge_updates / world / (ms_per_step * 1e-3)
        line["codewords_per_s"] = tf / (ms_per_step * 1e-3)
        line["fer"] = 1.0 - stats[0][0] / tf
        line["mean_iters"] = stats[0][1] / tf
        line["llr_checksum"] = {"rank0_sum_of_int32_words": int(parts[0].inp.view(torch.int32).sum(dtype=torch.int64).item())}
        ncu = read_ncu()
        lane_peak = 148 * 64 * f_max                    # ALU pipe: 16 lanes per SM sub-partition per clock (profiles/r01_pipe_ubench.md)
        issue_peak = 148 * 4 * f_max                    # one warp-instruction per SM sub-partition per clock
        alu = {"edge_pass_updates_per_s": upd_per_gpu_s, "peak": lane_peak, "unit": "lane-op/s",
               "peak_source": "148 SMs x 64 ALU-pipe lanes x sm_max_mhz",
               "note": "per GPU; one edge-pass update = one trip of either edge loop of src/decoder.rs:388-450"}
        issue = {"peak": issue_peak, "unit": "warp-instr/s", "peak_source": "148 SMs x 4 schedulers x sm_max_mhz"}
        if ncu:
            # ALU-pipe warp-instructions of the captured launch = busy fraction x 2 per clock per SM x active cycles x SMs
            alu_inst = ncu["alu_busy"] * 2.0 * ncu["sm_cycles"] * 148
            per_update = alu_inst / ncu["edge_updates"]
            achieved = per_update * upd_per_gpu_s * 32.0
            src = "profiles/r02_tm8192_ncu.csv (raw ncu export of %s) + profiles/r02_tm8192_ncu.json" % ncu["kernel"]
            alu.update({"achieved": achieved, "frac": achieved / lane_peak,
                        "alu_pipe_warp_instr_per_edge_update": per_update,
                        "alu_pipe_busy_frac_in_capture": ncu["alu_busy"], "source": src})
            # every warp-instruction of the captured launch takes one issue slot
            inst_per_update = ncu["inst_executed"] / ncu["edge_updates"]
            issue.update({"achieved": inst_per_update * upd_per_gpu_s, "frac": inst_per_update * upd_per_gpu_s / issue_peak,
                          "warp_instr_per_edge_update": inst_per_update, "issue_active_frac_in_capture": ncu["issue_active"],
                          "fma_heavy_pipe_busy_frac_in_capture": ncu["fmaheavy_busy"], "source": src})
            roofline["traffic"] = ncu["dram_bytes"] / ncu["frames"] * parts[0].frames
            roofline["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the tracked capture (%d frames), "
                                          "per frame, times this launch's frames" % ncu["frames"])
        roofline = {"bound": "issue_slots", "achieved": issue.get("achieved"), "peak": issue_peak, "unit": "warp-instr/s",
                    "frac": issue.get("frac"), "traffic": roofline["traffic"], "issue": issue, "alu": alu,
                    "hbm": {k: roofline[k] for k in ("achieved", "peak", "unit", "frac", "peak_source")},
                    "traffic_source": roofline.get("traffic_source"), "kernel_of": roofline["kernel_of"],
                    "share_of_step": 1.0,
                    "note": "decode_ms is an instruction-throughput kernel, not an HBM one (DESIGN.md 4.1).  Since the "
                            "self-correction rule and the sign of u moved to fp16 lanes the work is spread over the ALU pipe "
                            "and the FMA-heavy pipe, and the most utilised resource is the issue slots: the primary fraction "
                            "is warp-instructions issued against 148 SMs x 4 schedulers x f_max; the ALU-pipe fraction "
                            "(the primary one until the fp16 rewrite) and the HBM figure are secondary"}
    line["roofline"] = roofline

    if not args.no_cpu_baseline:
        oracle, okind = load_oracle()
        cores = len(os.sched_getaffinity(0))
        budget = (14.0 if world == 1 else 3.0) / len(parts)
        rates, compared, mism, samples = [], 0, 0, []
        for pt in parts:
            cap = 1 << 16 if pt.p["op"] != "enc" else 1 << 20
            host_inp = pt.inp[: min(pt.frames, cap)].cpu().numpy()
            cf, cs, cout = timed_oracle(oracle, pt, host_inp, cores, budget)
            rates.append(cf / cs)
            samples.append(cf)
            g_out = pt.out[:cf].cpu().numpy()
            if pt.p["op"] == "enc":
                bad = ~(g_out == cout[0]).all(axis=1)
            else:
                g_ok, g_it = pt.ok[:cf].cpu().numpy(), pt.iters[:cf].cpu().numpy()
                same = (g_out == cout[0]).all(axis=1) & (g_ok.astype(bool) == cout[1].astype(bool)) & \
                       (g_it.astype(np.int64) == cout[2].astype(np.int64))
                bad = ~same
            compared += cf
            mism += int(bad.sum())
        k_of = {p["code"]: pt.c.k() for p, pt in zip(spec, parts)}
        if world == 1:
            line["cpu_baseline"] = {"value": mix_rate(wl, spec, rates, k_of), "unit": unit, "cores": cores, "kind": "port",
                                    "codewords_per_s_by_part": rates,
                                    "sample": "first %s frames of rank 0's part(s) (the same input bytes), oracle build: %s; "
                                              "value = the workload's mix at the measured per-part rates" % (samples, okind)}
        float_parts = [p["ty"] in ("f32", "f64") and p["op"] == "ms" for p in spec]
        line["parity"] = {"frames_compared": compared, "mismatches": mism,
                          "checked": "decoded bytes, success flag, iteration count (decode) / codeword bytes (encode) vs CPU oracle",
                          "tolerance": "exact" if not any(float_parts) else
                                       "exact for integer parts; float parts: identical on >= 99.99 % of frames (north_star)"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
