"""ctypes binding of libholo_b200.so -- the only way the Python host reaches the CUDA kernels.

The prototypes are read from ``include/holo_b200.h`` so that the header is the single source of truth for the
C-ABI.  There is NO fallback: if the library is missing the import fails loudly (build it with
``python holo_diffusion_b200/build.py`` or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libholo_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "holo_b200.h")

_SCALARS = {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong, "double": ctypes.c_double}


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[object, List[object], List[str]]]:
    """name -> (restype, argtypes, argnames) for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int|void)\s+(holo_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = {"const char*": ctypes.c_char_p, "long long": ctypes.c_longlong, "int": ctypes.c_int, "void": None}[ret]
        argtypes, argnames = [], []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                nm = re.search(r"(\w+)$", a).group(1)
                ty = a[: -len(nm)].strip()
                if "*" in ty:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_SCALARS[ty.replace("const ", "").strip()])
                argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


class HoloError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python holo_diffusion_b200/build.py` "
                "(there is no CPU fallback).")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argtypes, _) in self.protos.items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if self.cdll.holo_version() != self._header_version():
            raise ImportError("libholo_b200.so is stale (version mismatch with include/holo_b200.h); rebuild")

    @staticmethod
    def _header_version() -> int:
        return int(re.search(r"#define HOLO_B200_VERSION (\d+)", open(HEADER_PATH).read()).group(1))

    def call(self, name: str, *args):
        """Call an int-returning entry point; raise HoloError with the library's message on failure."""
        rc = getattr(self.cdll, name)(*args)
        if rc != 0:
            msg = self.cdll.holo_last_error()
            raise HoloError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")

    def try_call(self, name: str, *args) -> int:
        return getattr(self.cdll, name)(*args)


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
