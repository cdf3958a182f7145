"""Host-side checks of the measurement harness (no GPU): the tracked ncu export that bench.py's roofline is recomputed
from parses and is self-consistent, the workload table covers the five BASELINE.json configurations, and the reference
arm (the oracle alone on the host cores) prints the contract's JSON line."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_tracked_ncu_export_parses_and_is_consistent():
    b = load_bench()
    ncu = b.read_ncu()
    assert ncu is not None, "profiles/r02_tm8192_ncu.csv / .json missing or unreadable"
    assert "decode_ms_tm_i8_kernel" in ncu["kernel"]
    assert ncu["frames"] == 65536 and ncu["edge_updates"] > 0
    for k in ("alu_busy", "issue_active", "fmaheavy_busy"):
        assert 0.0 < ncu[k] < 1.0, k
    # warp-instructions per issue slot: executed / (active cycles x 4 schedulers x 148 SMs) must equal the issue-active fraction
    frac = ncu["inst_executed"] / (ncu["sm_cycles"] * 4 * 148)
    assert abs(frac - ncu["issue_active"]) < 0.02, (frac, ncu["issue_active"])
    # DRAM traffic per frame stays at the algorithmic bytes (n + (n + p) / 8 + 8 for TM8192 i8: 9480), no wasted re-reads
    assert 0.9 * 9480 < ncu["dram_bytes"] / ncu["frames"] < 1.1 * 9480
    meta = json.load(open(os.path.join(ROOT, "profiles", "r02_tm8192_ncu.json")))
    sass = open(os.path.join(ROOT, "profiles", "r02_tm8192.sass")).readline()
    assert "decode_ms_tm_i8_kernel" in sass and meta["code"] == "TM8192"


def test_workload_table_covers_the_five_configurations():
    b = load_bench()
    ops = {}
    for w in ("c1", "c2", "c3", "c4", "c5"):
        parts = b.workload_parts(w)
        assert parts, w
        ops[w] = [(p["op"], p["code"], p.get("ty")) for p in parts]
    assert all(op == "ms" and code == 0 and ty == "i8" for op, code, ty in ops["c1"]) and len(ops["c1"]) == 5
    assert {(op, ty) for op, _, ty in ops["c2"]} >= {("ms", "i16"), ("ms", "f32")} and any(op == "bf" for op, _, _ in ops["c2"])
    assert ops["c3"] == [("ms", 8, "i8")]
    assert sorted(code for _, code, _ in ops["c4"]) == [6, 7]
    assert sorted(code for _, code, _ in ops["c5"]) == list(range(9)) and all(op == "enc" for op, _, _ in ops["c5"])


def test_reference_arm_prints_the_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0", "--frames-per-gpu", "64"], text=True, timeout=600,
                                  env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "Gbit/s"
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["value"] > 0 and line["config"]["workload"]
