"""ctypes binding of the C-ABI library (include/tatt_b200.h is the single source of truth:
prototypes are parsed from it, so every declared symbol must be exported by the .so).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "tatt_b200.h")
LIB_PATH = os.environ.get("TATT_LIB") or os.path.join(HERE, "lib", "libtatt_b200.so")   # TATT_LIB: kernel-variant experiments

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
    "float": ctypes.c_float,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """-> {name: (return type, [arg types])} for every prototype in the header."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(tatt_\w+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    types.append("ptr")
                else:
                    t = " ".join(a.split(" ")[:-1])
                    if t not in _CTYPES:
                        raise RuntimeError("unknown C type %r in %s" % (a, name))
                    types.append(t)
        protos[name] = (ret, types)
    return protos


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "tatt_b200: C-ABI library %s not built. Run `python -m tatt_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / PyTorch fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (ret, types) in _protos.items():
        try:
            fn = getattr(L, name)
        except AttributeError as e:
            raise RuntimeError("tatt_b200: symbol %s declared in the header is not exported" % name) from e
        fn.restype = ctypes.c_char_p if ret != "int" else ctypes.c_int
        fn.argtypes = [ctypes.c_void_p if t == "ptr" else _CTYPES[t] for t in types]
    _lib = L
    return L


def protos():
    lib()
    return _protos


def last_error() -> str:
    e = lib().tatt_last_error()
    return e.decode() if e else ""


# kernels launched per entry-point call (everything not listed launches exactly one)
_KERNELS = {"tatt_bn_stats": 2, "tatt_bn_bwd": 3, "tatt_memcpy_d2d": 0, "tatt_memset0": 0,
            # GEMM / conv entry points split their fp32 operands into bf16 planes first (2 passes unless the caller
            # passes pre-split planes); counted as the common case
            "tatt_gemm": 3, "tatt_conv2d_igemm": 3, "tatt_conv2d_wgrad": 3, "tatt_rows_wgrad": 2}
launch_count = 0


_profile = None          # list of (name, args, start event, end event) while a Profile is active


class Profile:
    """Per-entry-point device timing of everything launched inside the `with` block: a CUDA event pair on the launching
    (current) stream around every C-ABI call.  For eager (non-graph) steps; `summary()` groups by entry point and a
    caller-supplied shape key.  bench.py uses it to name the time-dominant kernel of a step live."""

    def __enter__(self):
        global _profile
        self.records = _profile = []
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None
        import torch
        torch.cuda.synchronize()
        self.rows = [(n, a, e0.elapsed_time(e1) * 1e-3) for n, a, e0, e1 in self.records]
        return False

    def summary(self, keyfn=None):
        """-> [(key, calls, total seconds)] sorted by total time; keyfn(name, args) -> hashable key."""
        agg = {}
        for n, a, t in self.rows:
            k = keyfn(n, a) if keyfn else n
            c = agg.setdefault(k, [0, 0.0, n, a])
            c[0] += 1
            c[1] += t
        return sorted(((k, v[0], v[1], v[2], v[3]) for k, v in agg.items()), key=lambda r: -r[2])


def call_host(name: str, *args):
    """Entry points that launch nothing (host-side helpers)."""
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, last_error()))


def call(name: str, *args):
    """Invoke an int-returning entry point; non-zero -> RuntimeError(tatt_last_error())."""
    global launch_count
    launch_count += _KERNELS.get(name, 1)
    if _profile is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib(), name)(*args)
        e1.record()
        _profile.append((name, args, e0, e1))
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, last_error()))
