"""ctypes binding of librbm_b200.so (the C ABI declared in include/rbm_b200.h).

This is the same binding a Julia host performs with `ccall`.  There is
no CPU fallback: if the shared library is missing, or no sm_100 GPU is present,
everything here raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)                       # reducedbasismethods.jl_b200/
LIB_PATH = os.environ.get("RBM_B200_LIB") or os.path.join(ROOT, "lib", "librbm_b200.so")   # same override as the Julia glue
CSRC = os.path.join(ROOT, "csrc")

EXPORTS = [
    "rbm_version", "rbm_init", "rbm_finalize", "rbm_last_error", "rbm_set_stream", "rbm_synchronize",
    "rbm_device_info", "rbm_device_malloc", "rbm_device_free", "rbm_memcpy", "rbm_host_register", "rbm_host_unregister",
    "rbm_launch_count",
    "rbm_profile_enable", "rbm_profile_query",
    "rbm_comm_unique_id", "rbm_comm_init", "rbm_comm_destroy", "rbm_comm_peer_handle", "rbm_comm_peer_attach",
    "rbm_pic_create", "rbm_pic_destroy", "rbm_pic_set_particles", "rbm_pic_set_conventions", "rbm_pic_run",
    "rbm_field_update", "rbm_field_eval", "rbm_field_energy", "rbm_field_coefficients",
    "rbm_locate", "rbm_deposit", "rbm_pic_step",
    "rbm_gram", "rbm_pod_evd", "rbm_pod_factor", "rbm_pod_basis", "rbm_evd_destroy", "rbm_deim", "rbm_project",
    "rbm_reduced_create", "rbm_reduced_destroy", "rbm_reduced_run",
    "rbm_sample_bump_on_tail", "rbm_growth_rate",
]

ERRORS = {-1: "RBM_ERR_INVALID", -2: "RBM_ERR_CUDA", -3: "RBM_ERR_NOMEM", -4: "RBM_ERR_UNSUPPORTED",
          -5: "RBM_ERR_NUMERIC", -6: "RBM_ERR_COMM"}


class RBMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None
_vp = C.c_void_p
_i64 = C.c_int64


def load():
    """dlopen the library and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C {CSRC}`.  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.rbm_version.restype = C.c_int
    lib.rbm_init.argtypes = [C.c_int, C.POINTER(_vp)]
    lib.rbm_finalize.argtypes = [_vp]
    lib.rbm_finalize.restype = None
    lib.rbm_last_error.argtypes = [_vp]
    lib.rbm_last_error.restype = C.c_char_p
    lib.rbm_set_stream.argtypes = [_vp, _vp]
    lib.rbm_synchronize.argtypes = [_vp]
    lib.rbm_device_info.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(_i64), C.POINTER(_i64)]
    lib.rbm_device_malloc.argtypes = [_vp, _i64, C.POINTER(_vp)]
    lib.rbm_device_free.argtypes = [_vp, _vp]
    lib.rbm_memcpy.argtypes = [_vp, _vp, _vp, _i64]
    lib.rbm_host_register.argtypes = [_vp, _vp, _i64]
    lib.rbm_host_unregister.argtypes = [_vp, _vp]
    lib.rbm_launch_count.argtypes = [_vp]
    lib.rbm_launch_count.restype = _i64
    lib.rbm_profile_enable.argtypes = [_vp, C.c_int]
    lib.rbm_profile_query.argtypes = [_vp, C.c_char_p, C.POINTER(_i64), C.POINTER(C.c_double)]
    lib.rbm_comm_unique_id.argtypes = [_vp]
    lib.rbm_comm_init.argtypes = [_vp, C.c_int, C.c_int, _vp]
    lib.rbm_comm_destroy.argtypes = [_vp]
    lib.rbm_comm_peer_handle.argtypes = [_vp, _vp]
    lib.rbm_comm_peer_attach.argtypes = [_vp, _vp]
    lib.rbm_pic_create.argtypes = [_vp, _i64, C.c_int, C.c_int, C.c_double, C.POINTER(_vp)]
    lib.rbm_pic_destroy.argtypes = [_vp]
    lib.rbm_pic_destroy.restype = None
    lib.rbm_pic_set_particles.argtypes = [_vp, _vp, _vp, _vp, C.c_double]
    lib.rbm_pic_set_conventions.argtypes = [_vp, C.c_double, C.c_int, C.c_double, C.c_int]
    lib.rbm_pic_run.argtypes = [_vp, _vp, C.c_int, C.c_double, C.c_int, C.c_int] + [_vp] * 7
    lib.rbm_field_update.argtypes = [_vp, _vp, _vp, C.c_double, _i64, C.c_double]
    lib.rbm_field_eval.argtypes = [_vp, _vp, _i64, _vp]
    lib.rbm_field_energy.argtypes = [_vp, C.POINTER(C.c_double)]
    lib.rbm_field_coefficients.argtypes = [_vp, _vp]
    lib.rbm_locate.argtypes = [_vp, _vp, _i64, _vp, _vp]
    lib.rbm_deposit.argtypes = [_vp, _vp, _vp, C.c_double, _i64, _vp]
    lib.rbm_pic_step.argtypes = [_vp, _vp, _vp, _vp, C.c_double, _i64, C.c_double, C.c_double, _vp, _vp, _vp]
    lib.rbm_gram.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _i64, _vp]
    lib.rbm_pod_evd.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _i64, C.c_int, C.c_double, _vp,
                                C.POINTER(C.c_int), _vp, _i64, C.c_int]
    lib.rbm_pod_factor.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _i64, C.c_int, C.c_double, _vp,
                                   C.POINTER(C.c_int), C.POINTER(_vp)]
    lib.rbm_pod_basis.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _i64, C.c_int, _vp, _i64]
    lib.rbm_evd_destroy.argtypes = [_vp]
    lib.rbm_evd_destroy.restype = None
    lib.rbm_deim.argtypes = [_vp, _vp, _i64, _i64, C.c_int, _vp]
    lib.rbm_project.argtypes = [_vp, _vp, _i64, _i64, C.c_int, _vp, _i64, _i64, _vp, _i64]
    lib.rbm_reduced_create.argtypes = [_vp, _vp, _i64, C.c_int, _vp, _i64, C.c_int, _vp, C.POINTER(_vp)]
    lib.rbm_reduced_destroy.argtypes = [_vp]
    lib.rbm_reduced_destroy.restype = None
    lib.rbm_reduced_run.argtypes = [_vp, _vp, C.c_int, C.c_double, C.c_int, C.c_int] + [_vp] * 8
    lib.rbm_sample_bump_on_tail.argtypes = [_vp, _i64, _i64, C.c_uint64] + [C.c_double] * 6 + [_vp, _vp, _vp,
                                                                                                C.POINTER(_i64)]
    lib.rbm_growth_rate.argtypes = [_vp, _vp, _vp, _i64, _i64, _i64, C.c_int, _vp, C.c_int, _vp, _vp, _vp]
    _lib = lib
    return lib


def ptr(a):
    """Raw address of a NumPy array (host) or a torch CUDA/CPU tensor; None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(f"cannot take the address of {type(a)}")


def nelem(a) -> int:
    if isinstance(a, np.ndarray):
        return int(a.size)
    return int(a.numel())


def is_contiguous(a) -> bool:
    if isinstance(a, np.ndarray):
        return bool(a.flags.c_contiguous or a.flags.f_contiguous)
    return bool(a.is_contiguous())


def check_buffer(a, n: int, name: str, writable: bool = True):
    """The C ABI sees only a raw address: refuse anything that is not a dense float64 buffer of exactly n elements
    (the reference would throw DimensionMismatch / BoundsError; the library would write past the end).  None passes."""
    if a is None or isinstance(a, int):
        return
    if not (isinstance(a, np.ndarray) or hasattr(a, "data_ptr")):
        raise TypeError(f"{name}: expected a NumPy array or a torch tensor, got {type(a)}")
    if not is_f64(a):
        raise ValueError(f"{name}: must be float64 (got {a.dtype})")
    if not is_contiguous(a):
        raise ValueError(f"{name}: must be contiguous (a strided view cannot be handed to the library)")
    if nelem(a) != int(n):
        raise ValueError(f"DimensionMismatch: {name} holds {nelem(a)} elements, expected {int(n)}")
    if writable and isinstance(a, np.ndarray) and not a.flags.writeable:
        raise ValueError(f"{name}: read-only array used as an output")


def is_f64(a) -> bool:
    if isinstance(a, np.ndarray):
        return a.dtype == np.float64
    if hasattr(a, "dtype"):
        return str(a.dtype) in ("torch.float64", "float64")
    return True


class Context:
    """rbm_ctx: one GPU, one stream, optional NCCL communicator."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _vp()
        rc = self.lib.rbm_init(int(device), C.byref(h))
        if rc != 0:
            raise RBMError(rc, self.lib.rbm_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, rc: int):
        if rc != 0:
            raise RBMError(rc, self.lib.rbm_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rbm_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        self.check(self.lib.rbm_set_stream(self.h, stream_ptr))

    def synchronize(self):
        self.check(self.lib.rbm_synchronize(self.h))

    def device_info(self):
        sm = C.c_int()
        fr, tot = _i64(), _i64()
        self.check(self.lib.rbm_device_info(self.h, C.byref(sm), C.byref(fr), C.byref(tot)))
        return sm.value, fr.value, tot.value

    def host_register(self, a):
        """Page-lock a host NumPy array (rbm_host_register): transfers to/from it then run at pinned speed."""
        self.check(self.lib.rbm_host_register(self.h, a.ctypes.data, a.nbytes))

    def host_unregister(self, a):
        self.check(self.lib.rbm_host_unregister(self.h, a.ctypes.data))

    def launch_count(self) -> int:
        return int(self.lib.rbm_launch_count(self.h))

    def profile_enable(self, on: bool):
        self.check(self.lib.rbm_profile_enable(self.h, 1 if on else 0))

    def profile_query(self, kernel: str):
        """-> (launches, total_ms) of the named kernel since profile_enable(True)."""
        n, ms = _i64(), C.c_double()
        self.check(self.lib.rbm_profile_query(self.h, kernel.encode(), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    # -- communicator -------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self.lib.rbm_comm_unique_id(buf)
        if rc != 0:
            raise RBMError(rc, self.lib.rbm_last_error(None).decode())
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self.check(self.lib.rbm_comm_init(self.h, nranks, rank, buf))

    def comm_peer_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.check(self.lib.rbm_comm_peer_handle(self.h, buf))
        return buf.raw

    def comm_peer_attach(self, handles: bytes | None):
        """handles: nranks x 64 bytes in rank order; None switches the peer-memory path off."""
        if handles is None:
            self.check(self.lib.rbm_comm_peer_attach(self.h, None))
            return
        buf = C.create_string_buffer(handles, len(handles))
        self.check(self.lib.rbm_comm_peer_attach(self.h, buf))

    def comm_destroy(self):
        self.check(self.lib.rbm_comm_destroy(self.h))
