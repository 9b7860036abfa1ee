"""ctypes binding of libddmp_b200.so (the C ABI declared in include/ddmp_b200.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the declared ABI.  There is no
fallback: if the shared library is missing, or a call fails (for instance no CUDA device), a ``RuntimeError`` is
raised — the product never routes around its CUDA kernels.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DDMP_LIB_PATH: load another build of the SAME library (kernel A/B experiments: scripts/build_variant.sh); never a fallback
LIB_PATH = os.environ.get("DDMP_LIB_PATH") or os.path.join(_HERE, "lib", "libddmp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ddmp_b200.h")

_CTYPES = {
    "int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
    "double": ctypes.c_double,
}


def parse_header(path: str = HEADER_PATH):
    """[(name, restype, [argtypes])] for every function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for m in re.finditer(r"^\s*(const char\*|int64_t|int)\s+(ddmp_\w+)\s*\((.*?)\)\s*;", text, flags=re.S | re.M):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        restype = {"const char*": ctypes.c_char_p, "int64_t": ctypes.c_int64, "int": ctypes.c_int}[ret]
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.replace("const ", "").split()[0]
                    argtypes.append(_CTYPES[ty])
        out.append((name, restype, argtypes))
    return out


class _Lib:
    def __init__(self):
        self._dll = None
        self._lock = threading.Lock()

    def _load(self):
        with self._lock:
            if self._dll is not None:
                return self._dll
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C dual_dmp_b200/csrc`). dual_dmp_b200 has no CPU fallback.")
            dll = ctypes.CDLL(LIB_PATH)
            for name, restype, argtypes in parse_header():
                fn = getattr(dll, name)
                fn.restype = restype
                fn.argtypes = argtypes
            self._dll = dll
            return dll

    @property
    def dll(self):
        return self._dll if self._dll is not None else self._load()

    def call(self, name: str, *args):
        """Call an int-returning entry point; raise on a non-zero return code."""
        rc = getattr(self.dll, name)(*args)
        if rc != 0:
            msg = self.dll.ddmp_last_error()
            raise RuntimeError(f"{name} failed (rc={rc}): {msg.decode() if msg else ''}")

    def query(self, name: str, *args):
        """Call an entry point that returns a value (sizes, counts)."""
        return getattr(self.dll, name)(*args)


lib = _Lib()


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, dtype, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (dual_dmp_b200 has no CPU path), got device {t.device}")
    if t.dtype != dtype:
        raise RuntimeError(f"{what}: expected dtype {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def set_device(device) -> None:
    idx = device.index if isinstance(device, torch.device) else int(device)
    if idx is None:
        idx = torch.cuda.current_device()
    lib.call("ddmp_set_device", idx)
