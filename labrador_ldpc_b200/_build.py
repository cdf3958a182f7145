"""In-tree build of liblabrador_ldpc.so (nvcc, sm_100a only).

    python -m labrador_ldpc_b200._build [--force] [--verbose]

The shared object lands in labrador_ldpc_b200/lib/ (git-ignored, but it
travels to the GPU box with the repo snapshot).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "liblabrador_ldpc.so")

SOURCES = ["code_tables.cpp", "runtime.cu", "capi.cu", "decode_ms_generic.cu", "decode_ms_tm.cu", "decode_ms_tm_i16.cu", "decode_ms_tm_cluster.cu", "decode_ms_tm_wide.cu", "decode_ms_tc.cu", "decode_ms_tc_x2.cu", "decode_bf.cu", "decode_bf_tm.cu", "decode_bf_tc.cu",
           "encode.cu", "encode_tm.cu", "convert.cu", "channel.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # f32/f64 min-sum must stay plain IEEE add/sub
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-Xcudafe", "--diag_suppress=177",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    with open(os.path.join(os.path.dirname(HERE), "include", "labrador_ldpc.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def needs_build():
    stamp_file = LIB + ".stamp"
    if not (os.path.exists(LIB) and os.path.exists(stamp_file)):
        return True
    # On the GPU box there may be no nvcc worth invoking; trust a matching stamp.
    return open(stamp_file).read().strip() != _stamp()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.rsplit(".", 1)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    log = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            log.append("== %s\n%s%s" % (src, r.stdout, r.stderr))
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s%s" % (src, r.stdout, r.stderr))
            objs.append(obj)
    with open(os.path.join(OBJDIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s%s" % (r.stdout, r.stderr))
    with open(LIB + ".stamp", "w") as f:
        f.write(_stamp())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
