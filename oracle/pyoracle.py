"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package
(labrador_ldpc_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CODES = ["TC128", "TC256", "TC512", "TM1280", "TM1536", "TM2048", "TM5120", "TM6144", "TM8192",
         "TM20480", "TM24576", "TM32768"]      # the last three: k = 16384, decode only (oracle.cpp PARAMS)

_DT = {
    "i8": (np.int8, ctypes.c_int8),
    "i16": (np.int16, ctypes.c_int16),
    "i32": (np.int32, ctypes.c_int32),
    "f32": (np.float32, ctypes.c_float),
    "f64": (np.float64, ctypes.c_double),
}


def build(native=False):
    """Compile the oracle with oracle/Makefile; returns the .so path."""
    target = "native" if native else "all"
    subprocess.check_call(["make", "-s", "-C", HERE, target])
    return os.path.join(HERE, "_build", "liboracle_native.so" if native else "liboracle.so")


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    def __init__(self, native=False):
        name = "liboracle_native.so" if native else "liboracle.so"
        path = os.path.join(HERE, "_build", name)
        src = os.path.join(HERE, "oracle.cpp")
        if (not os.path.exists(path)) or os.path.getmtime(path) < os.path.getmtime(src):
            path = build(native)
        self.lib = ctypes.CDLL(path)
        self.native = native
        L = self.lib
        vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.oracle_code_param.argtypes = [ci, ci]
        L.oracle_edges.argtypes = [ci, vp, vp, vp]
        L.oracle_copy_encode.argtypes = [ci, vp, vp, ci]
        L.oracle_copy_encode_batch.argtypes = [ci, vp, vp, sz, ci]
        L.oracle_decode_bf.argtypes = [ci, vp, vp, vp, sz, vp]
        L.oracle_decode_bf_batch.argtypes = [ci, vp, vp, sz, sz, vp, vp, ci]
        L.oracle_decode_erasures.argtypes = [ci, vp, vp, sz, vp]
        for s in _DT:
            getattr(L, "oracle_decode_ms_" + s).argtypes = [ci, vp, vp, vp, vp, sz, vp]
            getattr(L, "oracle_decode_ms_%s_batch" % s).argtypes = [ci, vp, vp, sz, sz, vp, vp, ci]
            getattr(L, "oracle_hard_to_llrs_" + s).argtypes = [ci, vp, vp]
            getattr(L, "oracle_llrs_to_hard_" + s).argtypes = [ci, vp, vp]

    # -- parameters --------------------------------------------------------
    def param(self, code, which):
        return self.lib.oracle_code_param(code, which)

    def n(self, code): return self.param(code, 0)
    def k(self, code): return self.param(code, 1)
    def p(self, code): return self.param(code, 2)
    def m(self, code): return self.param(code, 3)
    def b(self, code): return self.param(code, 4)
    def edges_count(self, code): return self.param(code, 5)
    def bf_working_len(self, code): return self.param(code, 6)
    def ms_working_len(self, code): return self.param(code, 7)
    def ms_working_u8_len(self, code): return self.param(code, 8)
    def output_len(self, code): return self.param(code, 9)

    def edges(self, code):
        e = self.edges_count(code)
        checks = np.zeros(e, np.uint32)
        vars_ = np.zeros(e, np.uint32)
        crc = np.zeros(1, np.uint32)
        cnt = self.lib.oracle_edges(code, _ptr(checks), _ptr(vars_), _ptr(crc))
        return cnt, checks, vars_, int(crc[0])

    # -- encode ------------------------------------------------------------
    def copy_encode(self, code, data, word=8):
        data = np.ascontiguousarray(data, np.uint8)
        assert data.size == self.k(code) // 8
        cw = np.zeros(self.n(code) // 8, np.uint8)
        rc = self.lib.oracle_copy_encode(code, _ptr(data), _ptr(cw), word)
        assert rc == 0
        return cw

    def copy_encode_batch(self, code, data, nthreads=1):
        data = np.ascontiguousarray(data, np.uint8)
        batch = data.shape[0]
        assert data.shape[1] == self.k(code) // 8
        cw = np.zeros((batch, self.n(code) // 8), np.uint8)
        rc = self.lib.oracle_copy_encode_batch(code, _ptr(data), _ptr(cw), batch, nthreads)
        assert rc == 0
        return cw

    # -- decode ------------------------------------------------------------
    def decode_ms(self, code, llrs, maxiters, ty=None):
        ty = ty or {np.dtype(v[0]): k for k, v in _DT.items()}[llrs.dtype]
        npdt = _DT[ty][0]
        llrs = np.ascontiguousarray(llrs, npdt)
        assert llrs.size == self.n(code)
        out = np.full(self.output_len(code), 0xAA, np.uint8)
        working = np.full(self.ms_working_len(code), 1, npdt)       # dirty on purpose
        working_u8 = np.full(self.ms_working_u8_len(code), 0xFF, np.uint8)
        iters = ctypes.c_size_t(0)
        ok = getattr(self.lib, "oracle_decode_ms_" + ty)(
            code, _ptr(llrs), _ptr(out), _ptr(working), _ptr(working_u8), maxiters,
            ctypes.addressof(iters))
        return bool(ok), int(iters.value), out

    def decode_ms_batch(self, code, llrs, maxiters, ty=None, nthreads=1):
        ty = ty or {np.dtype(v[0]): k for k, v in _DT.items()}[llrs.dtype]
        npdt = _DT[ty][0]
        llrs = np.ascontiguousarray(llrs, npdt)
        batch = llrs.shape[0]
        assert llrs.shape[1] == self.n(code)
        out = np.zeros((batch, self.output_len(code)), np.uint8)
        success = np.zeros(batch, np.uint8)
        iters = np.zeros(batch, np.uint32)
        rc = getattr(self.lib, "oracle_decode_ms_%s_batch" % ty)(
            code, _ptr(llrs), _ptr(out), batch, maxiters, _ptr(success), _ptr(iters), nthreads)
        assert rc == 0
        return out, success, iters

    def decode_bf(self, code, hard, maxiters):
        hard = np.ascontiguousarray(hard, np.uint8)
        assert hard.size == self.n(code) // 8
        out = np.full(self.output_len(code), 0xAA, np.uint8)
        working = np.full(self.bf_working_len(code), 0xFF, np.uint8)
        iters = ctypes.c_size_t(0)
        ok = self.lib.oracle_decode_bf(code, _ptr(hard), _ptr(out), _ptr(working), maxiters,
                                       ctypes.addressof(iters))
        return bool(ok), int(iters.value), out

    def decode_bf_batch(self, code, hard, maxiters, nthreads=1):
        hard = np.ascontiguousarray(hard, np.uint8)
        batch = hard.shape[0]
        out = np.zeros((batch, self.output_len(code)), np.uint8)
        success = np.zeros(batch, np.uint8)
        iters = np.zeros(batch, np.uint32)
        rc = self.lib.oracle_decode_bf_batch(code, _ptr(hard), _ptr(out), batch, maxiters,
                                             _ptr(success), _ptr(iters), nthreads)
        assert rc == 0
        return out, success, iters

    def decode_erasures(self, code, codeword_full, maxiters):
        cw = np.ascontiguousarray(codeword_full, np.uint8).copy()
        assert cw.size == self.output_len(code)
        working = np.full(self.bf_working_len(code), 0xFF, np.uint8)
        iters = ctypes.c_size_t(0)
        ok = self.lib.oracle_decode_erasures(code, _ptr(cw), _ptr(working), maxiters,
                                             ctypes.addressof(iters))
        return bool(ok), int(iters.value), cw

    # -- converters --------------------------------------------------------
    def hard_to_llrs(self, code, hard, ty):
        hard = np.ascontiguousarray(hard, np.uint8)
        llrs = np.zeros(self.n(code), _DT[ty][0])
        getattr(self.lib, "oracle_hard_to_llrs_" + ty)(code, _ptr(hard), _ptr(llrs))
        return llrs

    def llrs_to_hard(self, code, llrs, ty=None):
        ty = ty or {np.dtype(v[0]): k for k, v in _DT.items()}[llrs.dtype]
        llrs = np.ascontiguousarray(llrs, _DT[ty][0])
        out = np.full(self.n(code) // 8, 0xAA, np.uint8)
        getattr(self.lib, "oracle_llrs_to_hard_" + ty)(code, _ptr(llrs), _ptr(out))
        return out
