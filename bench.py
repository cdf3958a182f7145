#!/usr/bin/env python3
"""Device-timed harness for the batched LDPC hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload: TM8192 (k=4096, r=1/2) decode_ms with i8 LLRs at Eb/N0 = 2 dB, max_iters 100
(BASELINE.json configs[2]).  Weak scaling: every GPU decodes the same-sized shard of
independent codewords (default 2 Mi frames = the per-GPU shard of the 16 Mi-frame config on
8 GPUs); there is no collective on the data path.  A "step" is one pass of the decoder over
the rank's whole shard, LLRs already resident in HBM (16 GiB per step, far larger than L2).

`value` is whole-job decoded information Gbit/s (all frames count; FER is reported beside
it).  `e2e` is the same metric through the C ABI with HOST (pinned) buffers, host<->device
copies inside the timed region.  `cpu_baseline` is the CPU oracle (C++ restatement of the
reference; the Rust crate cannot be built in this image) on the box's host cores over a
bounded prefix of the very same LLR bytes -- which doubles as the exact-parity check.

`--impl reference` times the reference's CPU algorithm (the oracle) alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODE = 8                     # TM8192
EBN0_DB = 2.0
MAX_ITERS = 100
N, K_INFO, OUT_LEN, EDGES = 8192, 4096, 1280, 30720
ALG_BYTES_PER_FRAME = N * 1 + OUT_LEN + 8          # LLRs in + packed output + flags
NCU_DRAM_BYTES_PER_FRAME = 9391                    # profiles/r01_tm8192_ncu.md, final build (615.47 MB / 65536 frames)
NCU_ALU_PIPE_BUSY = 0.875                          # sm__pipe_alu_cycles_active of the same capture
WORKLOAD = "TM8192 (k=4096, r=1/2) decode_ms i8 LLRs, Eb/N0 2 dB, max_iters 100"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


def generate_shard(L, torch, frames, seed, first_frame):
    """random data -> encode -> BPSK + AWGN -> i8 LLRs with this library's own kernels (csrc/channel.cu,
    csrc/encode.cu).  The generator is counter-based (Philox keyed by `seed`, counter = frame index), so rank r
    producing frames [r * frames, (r + 1) * frames) yields its slice of one logical run, whatever the world size.
    LLR = 2 y / sigma^2, quantised as clamp(round(4 LLR), -31, 31)."""
    c = L.LDPCCode(CODE)
    llrs = torch.empty((frames, N), dtype=torch.int8, device="cuda")
    sigma2 = 1.0 / (2.0 * (K_INFO / N) * 10.0 ** (EBN0_DB / 10.0))
    chunk = 65536
    data = torch.empty((chunk, K_INFO // 8), dtype=torch.uint8, device="cuda")
    cw = torch.empty((chunk, N // 8), dtype=torch.uint8, device="cuda")
    for f0 in range(0, frames, chunk):
        nf = min(chunk, frames - f0)
        c.random_data_batch(seed, first_frame + f0, data[:nf])
        c.copy_encode_batch(data[:nf], cw[:nf])
        c.awgn_batch(cw[:nf], sigma2 ** 0.5, 4.0 * 2.0 / sigma2, seed, first_frame + f0, "i8", limit=31, out=llrs[f0:f0 + nf])
    torch.cuda.synchronize()
    return llrs


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


def cpu_reference_rate(llrs_host, nthreads, target_seconds=12.0):
    """Times the CPU oracle (one decoder per host thread) on a bounded prefix; returns
    (frames, seconds, outputs) -- outputs are reused for the parity check."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    try:
        oracle = pyoracle.Oracle(native=True)
    except Exception:
        oracle = pyoracle.Oracle(native=False)
    probe = min(len(llrs_host), 16 * nthreads)
    t = time.perf_counter()
    oracle.decode_ms_batch(CODE, llrs_host[:probe], MAX_ITERS, nthreads=nthreads)
    dt = max(time.perf_counter() - t, 1e-3)
    frames = int(min(len(llrs_host), max(probe, target_seconds * probe / dt)))
    t = time.perf_counter()
    out = oracle.decode_ms_batch(CODE, llrs_host[:frames], MAX_ITERS, nthreads=nthreads)
    secs = time.perf_counter() - t
    return frames, secs, out, ("native" if oracle.native else "generic")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle) alone, all host threads."""
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pyoracle
    from frames import make_frames
    cores = len(os.sched_getaffinity(0))
    try:
        oracle = pyoracle.Oracle(native=True)
    except Exception:
        oracle = pyoracle.Oracle(native=False)
    gen = pyoracle.Oracle(native=False)
    sample = max(4 * cores, 64)
    _, _, llrs = make_frames(gen, CODE, sample, EBN0_DB, seed=1, ty="i8")
    oracle.decode_ms_batch(CODE, llrs, MAX_ITERS, nthreads=cores)      # cold start (thread creation, page faults)
    t = time.perf_counter()
    oracle.decode_ms_batch(CODE, llrs, MAX_ITERS, nthreads=cores)
    rate = sample / max(time.perf_counter() - t, 1e-3)
    # size a step to ~3 s of CPU work so warmup+steps stay within a few minutes
    per_step = int(max(sample, min(rate * 3.0, 1 << 16)))
    if per_step > sample:
        _, _, llrs = make_frames(gen, CODE, per_step, EBN0_DB, seed=1, ty="i8")
    for _ in range(args.warmup):
        oracle.decode_ms_batch(CODE, llrs, MAX_ITERS, nthreads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        out, ok, iters = oracle.decode_ms_batch(CODE, llrs, MAX_ITERS, nthreads=cores)
    secs = time.perf_counter() - t
    cw_s = per_step * args.steps / secs
    gbit = cw_s * K_INFO / 1e9
    line = {
        "impl": "reference", "metric": "decoded_info_gbit_per_s", "value": gbit, "unit": "Gbit/s",
        "codewords_per_s": cw_s, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": per_step,
                   "note": "C++ restatement of labrador-ldpc 1.2.1 decode_ms on host cores (Rust toolchain absent)"},
        "cpu_baseline": {"value": gbit, "unit": "Gbit/s", "codewords_per_s": cw_s, "cores": cores, "kind": "port",
                         "sample": "%d frames per step, %d steps, one decoder per host thread" % (per_step, args.steps)},
        "e2e": {"value": gbit, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fer": float(1.0 - ok.mean()), "mean_iters": float(iters.mean()),
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=1 << 21)
    ap.add_argument("--e2e-frames", type=int, default=1 << 17)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
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
    c = L.LDPCCode(CODE)
    frames = args.frames_per_gpu

    llrs = generate_shard(L, torch, frames, seed=1, first_frame=rank * frames)
    out = torch.empty((frames, OUT_LEN), dtype=torch.uint8, device="cuda")
    ok = torch.empty((frames,), dtype=torch.uint8, device="cuda")
    iters = torch.empty((frames,), dtype=torch.int32, device="cuda")

    def step():
        c.decode_ms_batch(llrs, MAX_ITERS, output=out, success=ok, iters=iters)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = L.kernel_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps

    stats = torch.stack([ok.float().sum(), iters.float().sum(),
                         torch.where(ok.bool(), iters.float() + 1.0, iters.float()).sum()]).double()
    if dist is not None:
        dist.all_reduce(stats)
    total_frames = frames * world
    fer = 1.0 - float(stats[0]) / total_frames
    mean_iters = float(stats[1]) / total_frames
    edge_updates = 2.0 * EDGES * float(stats[2])          # 2E(iters+1) or 2E*max_iters
    cw_s = total_frames / (ms_per_step * 1e-3)
    gbit = cw_s * K_INFO / 1e9

    # ---- end-to-end through the C ABI with pinned HOST buffers ----
    ef = min(args.e2e_frames, frames)
    h_llrs = torch.empty((ef, N), dtype=torch.int8).pin_memory()
    h_llrs.copy_(llrs[:ef])
    h_out = torch.empty((ef, OUT_LEN), dtype=torch.uint8).pin_memory()
    h_ok = torch.empty((ef,), dtype=torch.uint8).pin_memory()
    h_it = torch.empty((ef,), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    def e2e_step():
        c.decode_ms_batch(h_llrs, MAX_ITERS, output=h_out, success=h_ok, iters=h_it)   # host pointers -> streamed

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = 3
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_gbit = ef * world * K_INFO / e2e_s / 1e9
    e2e_match = bool(torch.equal(h_out, out[:ef].cpu()) and torch.equal(h_it, iters[:ef].cpu()))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    llr_sum = int(llrs.view(torch.int32).sum(dtype=torch.int64).item())
    peak, peak_src = read_peaks()
    kernel_ms = ms_per_step                                 # one decode kernel per step (plus an 8-byte memset)
    achieved = ALG_BYTES_PER_FRAME * frames / (kernel_ms * 1e-3) / 1e9
    int_peak = 148 * 128 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6
    line = {
        "metric": "decoded_info_gbit_per_s", "value": gbit, "unit": "Gbit/s", "codewords_per_s": cw_s,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_gpu": frames, "frames_total": total_frames,
                   "llr_bytes_per_gpu": frames * N, "l2": "inputs (%.1f GiB per step) far exceed the 126 MB L2" % (frames * N / 2 ** 30),
                   "kernel": c.decode_ms_kernel_name("i8"), "sharding": "independent codewords per rank, no collective"},
        "fer": fer, "mean_iters": mean_iters,
        "llr_checksum": {"rank0_sum_of_int32_words": llr_sum, "generator": "philox4x32-10 seed 1, frames rank*frames_per_gpu.. (csrc/channel.cu)"},
        "e2e": {"value": e2e_gbit, "unit": "Gbit/s", "frames_per_gpu": ef, "h2d_bytes_per_step": ef * N,
                "d2h_bytes_per_step": ef * (OUT_LEN + 1 + 4), "seconds_per_step": e2e_s,
                "matches_device_resident_run": e2e_match,
                "api": "labrador_ldpc_decode_ms_i8_batch with pinned host pointers (wall clock around the blocking call)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_FRAME * frames, "peak_source": peak_src,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                                       "(65536 frames, profiles/r01_tm8192_ncu.md) scaled to this launch's frame count",
                     "algorithmic_bytes_per_frame": ALG_BYTES_PER_FRAME,
                     "note": "decode_ms is bound by the integer ALU pipe and shared memory, not HBM (DESIGN.md); "
                             "see roofline_alu"},
        "roofline_alu": {"edge_pass_updates_per_s": edge_updates / world / (ms_per_step * 1e-3),
                         "int_lane_ops_peak_per_s": int_peak,
                         "alu_pipe_busy_frac_ncu": NCU_ALU_PIPE_BUSY,
                         "note": "per GPU; one edge-pass update = one trip of either edge loop of src/decoder.rs:388-450; "
                                 "the binding unit is the integer ALU pipe, 87.5 % busy in the ncu capture "
                                 "(profiles/r01_tm8192_ncu.md)"},
    }

    if numa is not None:
        line["e2e"]["host_cores_local_to_gpu"] = numa
    if not args.no_cpu_baseline:
        cores = len(os.sched_getaffinity(0))
        host_llrs = llrs[: min(frames, 1 << 16)].cpu().numpy()
        # the CPU baseline is an N=1 figure; multi-rank runs only keep a short parity sample
        cf, cs, cout, kind = cpu_reference_rate(host_llrs, cores, target_seconds=12.0 if world == 1 else 2.0)
        g_out = out[:cf].cpu().numpy()
        g_ok = ok[:cf].cpu().numpy()
        g_it = iters[:cf].cpu().numpy()
        mism = int((~((g_out == cout[0]).all(axis=1) & (g_ok.astype(bool) == cout[1].astype(bool)) &
                      (g_it.astype(np.int64) == cout[2].astype(np.int64)))).sum())
        if world == 1:
            line["cpu_baseline"] = {"value": cf / cs * K_INFO / 1e9, "unit": "Gbit/s", "codewords_per_s": cf / cs,
                                    "cores": cores, "kind": "port",
                                    "sample": "first %d frames of rank 0's shard (same LLR bytes), %.1f s, oracle build: %s"
                                              % (cf, cs, kind)}
        line["parity"] = {"frames_compared": cf, "mismatches": mism,
                          "checked": "decoded bytes, success flag, iteration count vs CPU oracle"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
