"""ctypes binding of libpcm_b200.so.  Prototypes are parsed from include/pcm_b200.h so the Python
side can never drift from the C ABI; every call checks the status code and raises RuntimeError
with pcm_last_error().  There is NO fallback: if the library is missing the import fails loudly."""
from __future__ import annotations

import ctypes
import os
import re

from . import _build

_HEADER = os.path.join(_build.REPO, "include", "pcm_b200.h")

_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double,
    "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t, "pcm_stream_t": ctypes.c_void_p,
}


def parse_header(path: str = _HEADER):
    """-> {name: (restype, [(argtype, argname), ...])} for every function the header declares."""
    with open(path) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    protos = {}
    for m in re.finditer(r"\b(const char\*|long long|int)\s+(pcm_\w+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    parsed.append((ctypes.c_void_p, a.split("*")[-1].strip()))
                else:
                    toks = a.split(" ")
                    ty = " ".join(t for t in toks[:-1] if t != "const")
                    parsed.append((_CTYPES[ty], toks[-1]))
        protos[name] = (ctypes.c_char_p if "char" in ret else ctypes.c_longlong if "long" in ret else ctypes.c_int, parsed)
    return protos


PROTOS = parse_header()


def library_path() -> str:
    """The in-tree library; PCM_B200_LIB points at another build of the same ABI (kernel experiments)."""
    return os.environ.get("PCM_B200_LIB") or _build.LIB_PATH


class _Lib:
    def __init__(self):
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the pcm_b200 kernels)")
        self._dll = ctypes.CDLL(path)
        self._fn = {}
        for name, (ret, args) in PROTOS.items():
            fn = getattr(self._dll, name)          # AttributeError => header/library drift
            fn.restype = ret
            fn.argtypes = [a for a, _ in args]
            self._fn[name] = fn
        self.launches = 0                           # C-ABI calls issued (bench `gpu_launches`)
        self.trace = None                           # list -> per-call CUDA-event timing (bench roofline pass)
        self.last_call = None                       # name of the most recent C-ABI call (tests: which kernel ran)

    def last_error(self) -> str:
        return self._fn["pcm_last_error"]().decode()

    def version(self) -> int:
        return int(self._fn["pcm_version"]())

    def call(self, name: str, *args):
        self.last_call = name
        if self.trace is not None:
            import torch
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = self._fn[name](*args)
            e1.record()
            self.trace.append((name, args, e0, e1))
        else:
            rc = self._fn[name](*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed (status {rc}): {self.last_error()}")
        self.launches += 1

    def host_call(self, name: str, *args) -> int:
        """Host-only entry points (Kaggle I/O): no kernel launch, not counted; returns the status code."""
        return int(self._fn[name](*args))


_LIB = None


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB
