"""labrador_ldpc_b200 -- host-side mirror of the reference's `LDPCCode` interface.

The reference API (methods on `enum LDPCCode`, reference src/codes/mod.rs:37-66,
src/encoder.rs:292-315, src/decoder.rs:93-116, 243, 347, 484, 498) is mirrored
by :class:`LDPCCode` with the same method names and argument meaning, plus the
`_batch` variants.  Every compute method goes through the C ABI of
``lib/liblabrador_ldpc.so`` (include/labrador_ldpc.h) and from there to the
sm_100a CUDA kernels.  There is no CPU path: importing works without a GPU
(size getters are host arithmetic) but compute calls raise :class:`LdpcError`
when no CUDA device is usable, and the import itself fails if the shared
library has not been built.

Buffers may be numpy arrays (host memory) or torch tensors (host or CUDA).
Torch is only plumbing here (device memory and streams).
"""
import ctypes
import enum
import os

import numpy as np

from . import _build

__all__ = ["LDPCCode", "LdpcError", "lib", "LLR_TYPES", "init", "shutdown", "kernel_launch_count",
           "decode_ms_mixed"]

_LIB_PATH = _build.LIB
if _build.needs_build() and os.environ.get("LABRADOR_LDPC_NO_REBUILD") != "1":
    # the shared object is missing (fresh checkout) or older than the sources: (re)build it with nvcc rather than
    # run without it or with a stale binary.  This is the in-tree build, not a fallback: if it fails the import fails.
    try:
        _build.build()
    except Exception as exc:   # pragma: no cover
        raise ImportError("labrador_ldpc_b200: liblabrador_ldpc.so is missing or stale and building it failed "
                          "(there is no CPU fallback): %s" % exc)
if not os.path.exists(_LIB_PATH):
    raise ImportError(
        "labrador_ldpc_b200: %s is missing -- build it with `python -m labrador_ldpc_b200._build` "
        "(there is no CPU fallback)" % _LIB_PATH)
lib = ctypes.CDLL(_LIB_PATH)

LLR_TYPES = {"i8": 0, "i16": 1, "i32": 2, "f32": 3, "f64": 4}
_NP_OF = {"i8": np.int8, "i16": np.int16, "i32": np.int32, "f32": np.float32, "f64": np.float64}
_TY_OF_NP = {np.dtype(v): k for k, v in _NP_OF.items()}


class LdpcError(RuntimeError):
    pass


_vp, _sz, _ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
for _name in ("code_n", "code_k", "bf_working_len", "ms_working_len", "ms_working_u8_len", "output_len"):
    _f = getattr(lib, "labrador_ldpc_" + _name)
    _f.argtypes = [_ci]
    _f.restype = _sz
lib.labrador_ldpc_last_error.restype = ctypes.c_char_p
lib.labrador_ldpc_version.restype = ctypes.c_char_p
lib.labrador_ldpc_decode_ms_kernel_name.restype = ctypes.c_char_p
lib.labrador_ldpc_decode_ms_kernel_name.argtypes = [_ci, _ci]
lib.labrador_ldpc_kernel_launch_count.restype = ctypes.c_ulonglong
lib.labrador_ldpc_edge_table_crc.restype = ctypes.c_uint32
lib.labrador_ldpc_edge_table_crc.argtypes = [_ci]
lib.labrador_ldpc_cuda_init.argtypes = [_vp, _ci]
lib.labrador_ldpc_alloc_pinned.restype = _vp
lib.labrador_ldpc_alloc_pinned.argtypes = [_sz]
lib.labrador_ldpc_free_pinned.argtypes = [_vp]
lib.labrador_ldpc_encode.argtypes = [_ci, _vp]
lib.labrador_ldpc_encode.restype = None
lib.labrador_ldpc_copy_encode.argtypes = [_ci, _vp, _vp]
lib.labrador_ldpc_copy_encode.restype = None
lib.labrador_ldpc_decode_bf.argtypes = [_ci, _vp, _vp, _vp, _sz, _vp]
lib.labrador_ldpc_decode_bf.restype = ctypes.c_bool
lib.labrador_ldpc_decode_bf_batch.argtypes = [_ci, _vp, _vp, _sz, _sz, _vp, _vp]
lib.labrador_ldpc_encode_batch.argtypes = [_ci, _vp, _sz]
lib.labrador_ldpc_copy_encode_batch.argtypes = [_ci, _vp, _vp, _sz]
for _t in LLR_TYPES:
    getattr(lib, "labrador_ldpc_decode_ms_" + _t).argtypes = [_ci, _vp, _vp, _vp, _vp, _sz, _vp]
    getattr(lib, "labrador_ldpc_decode_ms_" + _t).restype = ctypes.c_bool
    getattr(lib, "labrador_ldpc_decode_ms_%s_batch" % _t).argtypes = [_ci, _vp, _vp, _sz, _sz, _vp, _vp]
    getattr(lib, "labrador_ldpc_hard_to_llrs_" + _t).argtypes = [_ci, _vp, _vp]
    getattr(lib, "labrador_ldpc_hard_to_llrs_" + _t).restype = None
    getattr(lib, "labrador_ldpc_llrs_to_hard_" + _t).argtypes = [_ci, _vp, _vp]
    getattr(lib, "labrador_ldpc_llrs_to_hard_" + _t).restype = None
    getattr(lib, "labrador_ldpc_hard_to_llrs_%s_batch" % _t).argtypes = [_ci, _vp, _vp, _sz]
    getattr(lib, "labrador_ldpc_llrs_to_hard_%s_batch" % _t).argtypes = [_ci, _vp, _vp, _sz]
lib.labrador_ldpc_decode_ms_batch_async.argtypes = [_ci, _ci, _vp, _vp, _sz, _sz, _vp, _vp, _vp]
lib.labrador_ldpc_decode_bf_batch_async.argtypes = [_ci, _vp, _vp, _sz, _sz, _vp, _vp, _vp]
lib.labrador_ldpc_copy_encode_batch_async.argtypes = [_ci, _vp, _vp, _sz, _vp]
lib.labrador_ldpc_hard_to_llrs_batch_async.argtypes = [_ci, _ci, _vp, _vp, _sz, _vp]
lib.labrador_ldpc_llrs_to_hard_batch_async.argtypes = [_ci, _ci, _vp, _vp, _sz, _vp]
_cf = ctypes.c_float
lib.labrador_ldpc_decode_ms_i8_soft_batch.argtypes = [_ci, _vp, _cf, _ci, _vp, _sz, _sz, _vp, _vp]
lib.labrador_ldpc_decode_ms_i16_soft_batch.argtypes = [_ci, _vp, _cf, _ci, _vp, _sz, _sz, _vp, _vp]
lib.labrador_ldpc_decode_ms_i8_hard_batch.argtypes = [_ci, _vp, _vp, _sz, _sz, _vp, _vp]
lib.labrador_ldpc_decode_ms_front_batch_async.argtypes = [_ci, _ci, _ci, _vp, _cf, _ci, _vp, _sz, _sz, _vp, _vp, _vp]
lib.labrador_ldpc_quantise_i8_batch.argtypes = [_ci, _vp, _cf, _ci, _vp, _sz]
lib.labrador_ldpc_quantise_i16_batch.argtypes = [_ci, _vp, _cf, _ci, _vp, _sz]
lib.labrador_ldpc_quantise_batch_async.argtypes = [_ci, _ci, _vp, _cf, _ci, _vp, _sz, _vp]
lib.labrador_ldpc_copy_control_batch.argtypes = [_ci, _ci, _vp, _vp, _sz, _vp, _vp]
FRONT_NONE, FRONT_SOFT_F32, FRONT_HARD = 0, 1, 2
_u64 = ctypes.c_uint64
lib.labrador_ldpc_random_data_batch.argtypes = [_ci, _u64, _u64, _vp, _sz]
lib.labrador_ldpc_random_data_batch_async.argtypes = [_ci, _u64, _u64, _vp, _sz, _vp]
lib.labrador_ldpc_awgn_batch.argtypes = [_ci, _ci, _vp, _cf, _cf, _ci, _u64, _u64, _vp, _sz]
lib.labrador_ldpc_awgn_batch_async.argtypes = [_ci, _ci, _vp, _cf, _cf, _ci, _u64, _u64, _vp, _sz, _vp]
lib.labrador_ldpc_count_errors_batch.argtypes = [_ci, _vp, _vp, _vp, _sz]
lib.labrador_ldpc_count_errors_batch_async.argtypes = [_ci, _vp, _vp, _vp, _sz, _vp]


def _check(rc):
    if rc != 0:
        raise LdpcError("labrador_ldpc error %d: %s" % (rc, lib.labrador_ldpc_last_error().decode()))


def init(devices=None):
    """labrador_ldpc_cuda_init: pick the devices used for host-buffer batches."""
    if devices is None:
        _check(lib.labrador_ldpc_cuda_init(None, 0))
    else:
        arr = (ctypes.c_int * len(devices))(*devices)
        _check(lib.labrador_ldpc_cuda_init(ctypes.cast(arr, _vp), len(devices)))


def shutdown():
    lib.labrador_ldpc_cuda_shutdown()


def kernel_launch_count():
    return int(lib.labrador_ldpc_kernel_launch_count())


def version():
    return lib.labrador_ldpc_version().decode()


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous(), "buffers must be contiguous"
        return x.data_ptr()
    assert x.flags["C_CONTIGUOUS"], "buffers must be contiguous"
    return x.ctypes.data


def _nbytes(x):
    return x.numel() * x.element_size() if _is_torch(x) else x.nbytes


def _llr_type(x, ty=None):
    if ty is not None:
        return ty
    if _is_torch(x):
        import torch
        return {torch.int8: "i8", torch.int16: "i16", torch.int32: "i32",
                torch.float32: "f32", torch.float64: "f64"}[x.dtype]
    return _TY_OF_NP[x.dtype]


def _alloc_like(ref, shape, dtype_np):
    """Allocate an output buffer of the same kind (numpy / torch host / torch cuda) as `ref`."""
    if _is_torch(ref):
        import torch
        tdt = {np.uint8: torch.uint8, np.uint32: torch.int32, np.int8: torch.int8, np.int16: torch.int16,
               np.int32: torch.int32, np.float32: torch.float32, np.float64: torch.float64}[dtype_np]
        return torch.empty(shape, dtype=tdt, device=ref.device)
    return np.empty(shape, dtype_np)


def _current_stream(x):
    if _is_torch(x) and x.is_cuda:
        import torch
        return torch.cuda.current_stream(x.device).cuda_stream
    return None


class LDPCCode(enum.IntEnum):
    """Mirror of `enum LDPCCode` (reference src/codes/mod.rs:37-66), same discriminants."""
    TC128 = 0
    TC256 = 1
    TC512 = 2
    TM1280 = 3
    TM1536 = 4
    TM2048 = 5
    TM5120 = 6
    TM6144 = 7
    TM8192 = 8
    # extension: the k = 16384 codes whose parity-check constants the reference carries without supporting them
    # (src/lib.rs:81-83); see include/labrador_ldpc.h
    TM20480 = 9
    TM24576 = 10
    TM32768 = 11

    # ---- parameters (reference src/codes/mod.rs:367-409, src/decoder.rs:93-116) ----
    def n(self): return int(lib.labrador_ldpc_code_n(int(self)))
    def k(self): return int(lib.labrador_ldpc_code_k(int(self)))
    def output_len(self): return int(lib.labrador_ldpc_output_len(int(self)))
    def punctured_bits(self): return self.output_len() * 8 - self.n()
    def decode_bf_working_len(self): return int(lib.labrador_ldpc_bf_working_len(int(self)))
    def decode_ms_working_len(self): return int(lib.labrador_ldpc_ms_working_len(int(self)))
    def decode_ms_working_u8_len(self): return int(lib.labrador_ldpc_ms_working_u8_len(int(self)))
    def paritycheck_sum(self):
        return (self.decode_ms_working_len() - 3 * self.n() - 3 * self.punctured_bits() + 2 * self.k()) // 2

    # ---- single-codeword API: same arguments as the reference ----
    def encode(self, codeword):
        """LDPCCode::encode (src/encoder.rs:293): first k bits in, last n-k bits out, in place."""
        if _nbytes(codeword) * 8 != self.n():
            raise ValueError("codeword must be n bits long")
        if _is_torch(codeword) and codeword.is_cuda:
            _check(lib.labrador_ldpc_copy_encode_batch_async(int(self), None, _ptr(codeword), 1, _current_stream(codeword)))
        else:
            lib.labrador_ldpc_encode(int(self), _ptr(codeword))
        return codeword

    def copy_encode(self, data, codeword):
        """LDPCCode::copy_encode (src/encoder.rs:309)."""
        if _nbytes(data) * 8 != self.k():
            raise ValueError("data must be k bits long")
        if _nbytes(codeword) * 8 != self.n():
            raise ValueError("codeword must be n bits long")
        stream = _current_stream(codeword)
        if stream is not None:
            _check(lib.labrador_ldpc_copy_encode_batch_async(int(self), _ptr(data), _ptr(codeword), 1, stream))
        else:
            _check(lib.labrador_ldpc_copy_encode_batch(int(self), _ptr(data), _ptr(codeword), 1))
        return codeword

    def decode_bf(self, input, output, working=None, maxiters=50):
        """LDPCCode::decode_bf (src/decoder.rs:243) -> (success, iters)."""
        if _nbytes(input) != self.n() // 8:
            raise ValueError("input.len() != n/8")
        if _nbytes(output) != self.output_len():
            raise ValueError("output.len != (n+p)/8")
        if working is not None and _nbytes(working) != self.decode_bf_working_len():
            raise ValueError("working.len() incorrect")
        ok = np.zeros(1, np.uint8)
        it = np.zeros(1, np.uint32)
        if _is_torch(input) and input.is_cuda:
            okd = _alloc_like(input, (1,), np.uint8)
            itd = _alloc_like(input, (1,), np.uint32)
            _check(lib.labrador_ldpc_decode_bf_batch_async(int(self), _ptr(input), _ptr(output), 1, maxiters,
                                                           _ptr(okd), _ptr(itd), _current_stream(input)))
            return bool(okd.item()), int(itd.item())      # .item() synchronises with the stream the kernel ran on
        _check(lib.labrador_ldpc_decode_bf_batch(int(self), _ptr(input), _ptr(output), 1, maxiters,
                                                 _ptr(ok), _ptr(it)))
        return bool(ok[0]), int(it[0])

    def decode_ms(self, llrs, output, working=None, working_u8=None, maxiters=50, ty=None):
        """LDPCCode::decode_ms<T> (src/decoder.rs:347) -> (success, iters)."""
        ty = _llr_type(llrs, ty)
        esize = np.dtype(_NP_OF[ty]).itemsize
        if _nbytes(llrs) != self.n() * esize:
            raise ValueError("llrs.len() != n")
        if _nbytes(output) != self.output_len():
            raise ValueError("output.len() != (n+p)/8")
        if working is not None and _nbytes(working) != self.decode_ms_working_len() * esize:
            raise ValueError("working.len() incorrect")
        if working_u8 is not None and _nbytes(working_u8) != self.decode_ms_working_u8_len():
            raise ValueError("working_u8 != (n+p-k)/8")
        fn = getattr(lib, "labrador_ldpc_decode_ms_%s_batch" % ty)
        if _is_torch(llrs) and llrs.is_cuda:
            okd = _alloc_like(llrs, (1,), np.uint8)
            itd = _alloc_like(llrs, (1,), np.uint32)
            _check(lib.labrador_ldpc_decode_ms_batch_async(int(self), LLR_TYPES[ty], _ptr(llrs), _ptr(output), 1, maxiters,
                                                           _ptr(okd), _ptr(itd), _current_stream(llrs)))
            return bool(okd.item()), int(itd.item())
        ok = np.zeros(1, np.uint8)
        it = np.zeros(1, np.uint32)
        _check(fn(int(self), _ptr(llrs), _ptr(output), 1, maxiters, _ptr(ok), _ptr(it)))
        return bool(ok[0]), int(it[0])

    def hard_to_llrs(self, input, llrs, ty=None):
        """LDPCCode::hard_to_llrs<T> (src/decoder.rs:484)."""
        ty = _llr_type(llrs, ty)
        if _nbytes(input) != self.n() // 8:
            raise ValueError("input.len() != n/8")
        if _nbytes(llrs) != self.n() * np.dtype(_NP_OF[ty]).itemsize:
            raise ValueError("llrs.len() != n")
        stream = _current_stream(llrs)
        if stream is not None:
            _check(lib.labrador_ldpc_hard_to_llrs_batch_async(int(self), LLR_TYPES[ty], _ptr(input), _ptr(llrs), 1, stream))
        else:
            _check(getattr(lib, "labrador_ldpc_hard_to_llrs_%s_batch" % ty)(int(self), _ptr(input), _ptr(llrs), 1))
        return llrs

    def llrs_to_hard(self, llrs, output, ty=None):
        """LDPCCode::llrs_to_hard<T> (src/decoder.rs:498)."""
        ty = _llr_type(llrs, ty)
        if _nbytes(llrs) != self.n() * np.dtype(_NP_OF[ty]).itemsize:
            raise ValueError("llrs.len() != n")
        if _nbytes(output) != self.n() // 8:
            raise ValueError("output.len() != n/8")
        stream = _current_stream(llrs)
        if stream is not None:
            _check(lib.labrador_ldpc_llrs_to_hard_batch_async(int(self), LLR_TYPES[ty], _ptr(llrs), _ptr(output), 1, stream))
        else:
            _check(getattr(lib, "labrador_ldpc_llrs_to_hard_%s_batch" % ty)(int(self), _ptr(llrs), _ptr(output), 1))
        return output

    # ---- batched API (frame-major [batch][len] buffers) ----
    def _batch_of(self, x, per_frame_bytes):
        nb = _nbytes(x)
        if per_frame_bytes == 0 or nb % per_frame_bytes:
            raise ValueError("buffer is not a whole number of frames")
        return nb // per_frame_bytes

    @staticmethod
    def _check_frames(name, buf, batch, per_frame_bytes):
        """The C side writes `batch` frames into every result array: a caller-supplied one must hold exactly that."""
        if buf is not None and _nbytes(buf) != batch * per_frame_bytes:
            raise ValueError("%s must hold %d frames of %d bytes (has %d bytes)" % (name, batch, per_frame_bytes, _nbytes(buf)))

    def copy_encode_batch(self, data, codewords=None, stream=None):
        batch = self._batch_of(data, self.k() // 8)
        if codewords is None:
            codewords = _alloc_like(data, (batch, self.n() // 8), np.uint8)
        if self._batch_of(codewords, self.n() // 8) != batch:
            raise ValueError("codewords has the wrong number of frames")
        stream = stream if stream is not None else _current_stream(data)
        if stream is not None:
            _check(lib.labrador_ldpc_copy_encode_batch_async(int(self), _ptr(data), _ptr(codewords), batch, stream))
        else:
            _check(lib.labrador_ldpc_copy_encode_batch(int(self), _ptr(data), _ptr(codewords), batch))
        return codewords

    def encode_batch(self, codewords):
        batch = self._batch_of(codewords, self.n() // 8)
        stream = _current_stream(codewords)
        if stream is not None:     # data == NULL: in place (csrc/capi.cu encode_impl)
            _check(lib.labrador_ldpc_copy_encode_batch_async(int(self), None, _ptr(codewords), batch, stream))
        else:
            _check(lib.labrador_ldpc_encode_batch(int(self), _ptr(codewords), batch))
        return codewords

    def decode_ms_batch(self, llrs, maxiters, output=None, success=None, iters=None, ty=None, stream=None):
        """Batched decode_ms -> (output[B, output_len], success[B] u8, iters[B] u32)."""
        ty = _llr_type(llrs, ty)
        batch = self._batch_of(llrs, self.n() * np.dtype(_NP_OF[ty]).itemsize)
        if output is None:
            output = _alloc_like(llrs, (batch, self.output_len()), np.uint8)
        if success is None:
            success = _alloc_like(llrs, (batch,), np.uint8)
        if iters is None:
            iters = _alloc_like(llrs, (batch,), np.uint32)
        self._check_frames("output", output, batch, self.output_len())
        self._check_frames("success", success, batch, 1)
        self._check_frames("iters", iters, batch, 4)
        stream = stream if stream is not None else _current_stream(llrs)
        if stream is not None:
            _check(lib.labrador_ldpc_decode_ms_batch_async(int(self), LLR_TYPES[ty], _ptr(llrs), _ptr(output), batch,
                                                           maxiters, _ptr(success), _ptr(iters), stream))
        else:
            _check(getattr(lib, "labrador_ldpc_decode_ms_%s_batch" % ty)(
                int(self), _ptr(llrs), _ptr(output), batch, maxiters, _ptr(success), _ptr(iters)))
        return output, success, iters

    # ---- fused front ends (include/labrador_ldpc.h "Fused front ends"; csrc/front.cuh) ----
    def _decode_front(self, front, ty, src, batch, scale, limit, maxiters, output, success, iters, stream):
        if output is None:
            output = _alloc_like(src, (batch, self.output_len()), np.uint8)
        if success is None:
            success = _alloc_like(src, (batch,), np.uint8)
        if iters is None:
            iters = _alloc_like(src, (batch,), np.uint32)
        self._check_frames("output", output, batch, self.output_len())
        self._check_frames("success", success, batch, 1)
        self._check_frames("iters", iters, batch, 4)
        stream = stream if stream is not None else _current_stream(src)
        if stream is not None:
            _check(lib.labrador_ldpc_decode_ms_front_batch_async(
                int(self), LLR_TYPES[ty], front, _ptr(src), scale, limit, _ptr(output), batch, maxiters,
                _ptr(success), _ptr(iters), stream))
        elif front == FRONT_HARD:
            _check(lib.labrador_ldpc_decode_ms_i8_hard_batch(int(self), _ptr(src), _ptr(output), batch, maxiters,
                                                             _ptr(success), _ptr(iters)))
        else:
            _check(getattr(lib, "labrador_ldpc_decode_ms_%s_soft_batch" % ty)(
                int(self), _ptr(src), scale, limit, _ptr(output), batch, maxiters, _ptr(success), _ptr(iters)))
        return output, success, iters

    def decode_ms_soft_batch(self, soft, scale, limit, maxiters, ty="i8", output=None, success=None, iters=None,
                             stream=None):
        """decode_ms::<ty> of clamp(rint(soft * scale), -limit, limit), quantised inside the decoder (f32 soft values)."""
        if ty not in ("i8", "i16"):
            raise ValueError("soft front end quantises to i8 or i16")
        if _llr_type(soft) != "f32":
            raise ValueError("soft values must be float32")
        batch = self._batch_of(soft, self.n() * 4)
        return self._decode_front(FRONT_SOFT_F32, ty, soft, batch, float(scale), int(limit), maxiters, output, success,
                                  iters, stream)

    def decode_ms_hard_batch(self, input, maxiters, output=None, success=None, iters=None, stream=None):
        """decode_ms::<i8> of hard_to_llrs(input), converted inside the decoder (bit-packed hard decisions)."""
        batch = self._batch_of(input, self.n() // 8)
        return self._decode_front(FRONT_HARD, "i8", input, batch, 1.0, 0, maxiters, output, success, iters, stream)

    def quantise_batch(self, soft, scale, limit, ty="i8", llrs=None, stream=None):
        """llrs = clamp(rint(soft * scale), -limit, limit) as i8 / i16 (the stand-alone form of the soft front end)."""
        if ty not in ("i8", "i16"):
            raise ValueError("quantise produces i8 or i16")
        batch = self._batch_of(soft, self.n() * 4)
        if llrs is None:
            llrs = _alloc_like(soft, (batch, self.n()), _NP_OF[ty])
        self._check_frames("llrs", llrs, batch, self.n() * np.dtype(_NP_OF[ty]).itemsize)
        stream = stream if stream is not None else _current_stream(soft)
        if stream is not None:
            _check(lib.labrador_ldpc_quantise_batch_async(int(self), LLR_TYPES[ty], _ptr(soft), float(scale), int(limit),
                                                          _ptr(llrs), batch, stream))
        else:
            _check(getattr(lib, "labrador_ldpc_quantise_%s_batch" % ty)(int(self), _ptr(soft), float(scale), int(limit),
                                                                        _ptr(llrs), batch))
        return llrs

    # ---- harness kernels (include/labrador_ldpc.h "Harness kernels"; csrc/channel.cu) ----
    def random_data_batch(self, seed, first_frame, data):
        """Fill data[B, k/8] with the Philox bytes of frames first_frame .. first_frame+B-1 of run `seed`."""
        batch = self._batch_of(data, self.k() // 8)
        stream = _current_stream(data)
        if stream is not None:
            _check(lib.labrador_ldpc_random_data_batch_async(int(self), seed, first_frame, _ptr(data), batch, stream))
        else:
            _check(lib.labrador_ldpc_random_data_batch(int(self), seed, first_frame, _ptr(data), batch))
        return data

    def awgn_batch(self, codewords, sigma, scale, seed, first_frame, ty="f32", limit=0, out=None):
        """BPSK + Gaussian noise: y = (1 - 2 bit) + sigma z; out = y * scale (f32) or its quantisation (i8 / i16)."""
        batch = self._batch_of(codewords, self.n() // 8)
        if out is None:
            out = _alloc_like(codewords, (batch, self.n()), _NP_OF[ty])
        self._check_frames("out", out, batch, self.n() * np.dtype(_NP_OF[ty]).itemsize)
        stream = _current_stream(codewords)
        if stream is not None:
            _check(lib.labrador_ldpc_awgn_batch_async(int(self), LLR_TYPES[ty], _ptr(codewords), float(sigma), float(scale),
                                                      int(limit), seed, first_frame, _ptr(out), batch, stream))
        else:
            _check(lib.labrador_ldpc_awgn_batch(int(self), LLR_TYPES[ty], _ptr(codewords), float(sigma), float(scale),
                                                int(limit), seed, first_frame, _ptr(out), batch))
        return out

    def count_errors_batch(self, decoded, data, errors=None):
        """errors[f] = number of wrong bits among the first k/8 bytes of decoded[f] (decoded is [B, output_len])."""
        batch = self._batch_of(data, self.k() // 8)
        if self._batch_of(decoded, self.output_len()) != batch:
            raise ValueError("decoded has the wrong number of frames")
        if errors is None:
            errors = _alloc_like(data, (batch,), np.uint32)
        self._check_frames("errors", errors, batch, 4)
        stream = _current_stream(data)
        if stream is not None:
            _check(lib.labrador_ldpc_count_errors_batch_async(int(self), _ptr(decoded), _ptr(data), _ptr(errors), batch, stream))
        else:
            _check(lib.labrador_ldpc_count_errors_batch(int(self), _ptr(decoded), _ptr(data), _ptr(errors), batch))
        return errors

    def decode_bf_batch(self, input, maxiters, output=None, success=None, iters=None, stream=None):
        batch = self._batch_of(input, self.n() // 8)
        if output is None:
            output = _alloc_like(input, (batch, self.output_len()), np.uint8)
        if success is None:
            success = _alloc_like(input, (batch,), np.uint8)
        if iters is None:
            iters = _alloc_like(input, (batch,), np.uint32)
        self._check_frames("output", output, batch, self.output_len())
        self._check_frames("success", success, batch, 1)
        self._check_frames("iters", iters, batch, 4)
        stream = stream if stream is not None else _current_stream(input)
        if stream is not None:
            _check(lib.labrador_ldpc_decode_bf_batch_async(int(self), _ptr(input), _ptr(output), batch, maxiters,
                                                           _ptr(success), _ptr(iters), stream))
        else:
            _check(lib.labrador_ldpc_decode_bf_batch(int(self), _ptr(input), _ptr(output), batch, maxiters,
                                                     _ptr(success), _ptr(iters)))
        return output, success, iters

    def hard_to_llrs_batch(self, input, ty, llrs=None, stream=None):
        batch = self._batch_of(input, self.n() // 8)
        if llrs is None:
            llrs = _alloc_like(input, (batch, self.n()), _NP_OF[ty])
        self._check_frames("llrs", llrs, batch, self.n() * np.dtype(_NP_OF[ty]).itemsize)
        stream = stream if stream is not None else _current_stream(input)
        if stream is not None:
            _check(lib.labrador_ldpc_hard_to_llrs_batch_async(int(self), LLR_TYPES[ty], _ptr(input), _ptr(llrs),
                                                              batch, stream))
        else:
            _check(getattr(lib, "labrador_ldpc_hard_to_llrs_%s_batch" % ty)(int(self), _ptr(input), _ptr(llrs), batch))
        return llrs

    def llrs_to_hard_batch(self, llrs, output=None, ty=None, stream=None):
        ty = _llr_type(llrs, ty)
        batch = self._batch_of(llrs, self.n() * np.dtype(_NP_OF[ty]).itemsize)
        if output is None:
            output = _alloc_like(llrs, (batch, self.n() // 8), np.uint8)
        self._check_frames("output", output, batch, self.n() // 8)
        stream = stream if stream is not None else _current_stream(llrs)
        if stream is not None:
            _check(lib.labrador_ldpc_llrs_to_hard_batch_async(int(self), LLR_TYPES[ty], _ptr(llrs), _ptr(output),
                                                              batch, stream))
        else:
            _check(getattr(lib, "labrador_ldpc_llrs_to_hard_%s_batch" % ty)(int(self), _ptr(llrs), _ptr(output), batch))
        return output

    def copy_control_batch(self, llrs, output, success, iters, ty=None):
        """The host<->device transport of a host-buffer decode_ms_batch call without the kernel (bench.py's control)."""
        ty = _llr_type(llrs, ty)
        batch = self._batch_of(llrs, self.n() * np.dtype(_NP_OF[ty]).itemsize)
        self._check_frames("output", output, batch, self.output_len())
        self._check_frames("success", success, batch, 1)
        self._check_frames("iters", iters, batch, 4)
        _check(lib.labrador_ldpc_copy_control_batch(int(self), LLR_TYPES[ty], _ptr(llrs), _ptr(output), batch,
                                                    _ptr(success), _ptr(iters)))

    def decode_ms_kernel_name(self, ty):
        return lib.labrador_ldpc_decode_ms_kernel_name(int(self), LLR_TYPES[ty]).decode()

    def edge_table_crc(self):
        return int(lib.labrador_ldpc_edge_table_crc(int(self)))


def decode_ms_mixed(jobs, maxiters):
    """Mixed-code batch (BASELINE.json configs[3]): `jobs` is a list of (LDPCCode, llrs) with CUDA tensors.
    Every homogeneous sub-batch is enqueued on its own CUDA stream through the `_batch_async` C entry point
    (the kernels are specialised per code), and the caller's current stream waits for all of them.
    Returns a list of (output, success, iters) in job order."""
    import torch
    results = []
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in jobs]
    for (code, llrs), st in zip(jobs, streams):
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            results.append(code.decode_ms_batch(llrs, maxiters, stream=st.cuda_stream))
    for st in streams:
        cur.wait_stream(st)
    return results
