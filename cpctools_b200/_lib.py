"""ctypes binding of ``libsoapb200.so`` (the C ABI declared in ``include/soap_b200.h``).

The library is built IN-TREE by :func:`build_library` (called from ``__graft_entry__.build()``)
with ``nvcc -gencode arch=compute_100a,code=sm_100a`` and lives next to this file.  There is no
CPU fallback: if the library or a CUDA device is missing every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from ctypes import c_char_p, c_int, c_int64, c_size_t, c_uint64, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libsoapb200.so")
SOURCES = ("capi.cu", "expand.cu", "timelag.cu", "pairs.cu", "pairs_dmma.cu", "pairs_tf32.cu", "pairs_i8.cu", "classify.cu", "transitions.cu")
NVCC_FLAGS = (
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--threads", "0",
)

# dtype / kind / path codes of include/soap_b200.h
F32, F64 = 0, 1
KIND_SIMPLE, KIND_KERNEL, KIND_NORMALIZED, KIND_DOT, KIND_COSCLIP, KIND_EUCLID, KIND_NORMALIZED_FUSED, KIND_EUCLID_NORMALIZED = range(8)
PAIRS_AUTO, PAIRS_SIMT, PAIRS_TENSOR, PAIRS_TENSOR_NATIVE = 0, 1, 2, 3
MAX_WINDOWS = 4
NO_EXCLUDE = -(2 ** 63)  # SOAP_NO_EXCLUDE

# every symbol include/soap_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "soap_abi_version": (c_int, []),
    "soap_last_error": (c_char_p, []),
    "soap_launch_count": (c_int64, []),
    "soap_reserve_sms": (c_int, [c_int]),
    "soap_profile_enable": (c_int, [c_int]),
    "soap_profile_read": (c_int, [c_int, POINTER(ctypes.c_float), c_int]),
    "soap_expand": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "soap_expand_normalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "soap_timelag_last_kernel": (c_char_p, []),
    "soap_timelag_scratch_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, POINTER(c_int), c_int]),
    "soap_timelag": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64, c_int64, c_int,
                             POINTER(c_int), c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "soap_diff_transposed": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "soap_pairs_scratch_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "soap_pairs": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int, c_int,
                           c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "soap_pairs_rowmin_scratch_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "soap_pairs_rowmin": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int, c_int, c_int, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "soap_classify": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "soap_transition_counts": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "soap_residence_capacity": (c_int64, [c_int64, c_int64, c_int64, c_int64]),
    "soap_residence_events": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                      c_void_p, c_void_p]),
    "soap_copy_rows": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_void_p]),
    "soap_interleave_shards": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "soap_fill_uniform": (c_int, [c_void_p, c_int64, c_uint64, c_int, c_void_p]),
}


class SoapB200Error(RuntimeError):
    """An error reported by libsoapb200 (bad argument, CUDA failure, unsupported case)."""


def _needs_rebuild() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [
        os.path.join(CSRC, "common.cuh"),
        os.path.join(CSRC, "tc05.cuh"),
        os.path.join(CSRC, "timelag_rb.cuh"),
        os.path.join(CSRC, "timelag_pair.cuh"),
        os.path.join(CSRC, "timelag_chain.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "soap_b200.h"),
    ]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/*.cu`` into ``libsoapb200.so`` for sm_100a (cross-compiles without a GPU)."""
    if not force and not _needs_rebuild():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise SoapB200Error("nvcc not found: libsoapb200.so cannot be built")
    cmd = [nvcc, *NVCC_FLAGS, "-shared", "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(" ".join(cmd))
        print(proc.stdout, proc.stderr)
    if proc.returncode != 0:
        raise SoapB200Error("building libsoapb200.so failed:\n" + proc.stderr[-4000:])
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def load() -> ctypes.CDLL:
    """Load the library and bind every declared symbol.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("SOAP_B200_LIB", LIB_PATH)  # A/B benchmarking of two builds
        if not os.path.exists(path):
            raise SoapB200Error(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(cpctools_b200 has no CPU fallback)"
            )
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI and the header disagree
            fn.restype, fn.argtypes = res, args
        if lib.soap_abi_version() != 1:
            raise SoapB200Error("libsoapb200.so ABI version mismatch: rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().soap_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)
        raise SoapB200Error(f"{what}: {msg} (code {rc})")


PROF_EXPAND, PROF_TIMELAG, PROF_PAIRS = 0, 1, 2


def profile_enable(on: bool = True) -> None:
    check(load().soap_profile_enable(1 if on else 0), "soap_profile_enable")


def profile_read(kernel_id: int, capacity: int = 4096) -> "list[float]":
    """Elapsed ms (CUDA events on the launching stream) of the main-kernel launches since enable."""
    buf = (ctypes.c_float * capacity)()
    n = load().soap_profile_read(kernel_id, buf, capacity)
    if n < 0:
        check(n, "soap_profile_read")
    return [float(buf[i]) for i in range(min(n, capacity))]


def launch_count() -> int:
    return int(load().soap_launch_count())
