"""ctypes binding of libdcue_b200.so (the C ABI declared in include/dcue_b200.h).

The prototypes are parsed from the header itself, so the Python side cannot drift from the C
ABI.  There is no CPU fallback: if the library is missing or a kernel fails, an exception is
raised.  Build with ``python __graft_entry__.py`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdcue_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dcue_b200.h")

FMT_F16, FMT_BF16 = 0, 1
IMPL_SIMT, IMPL_TC = 0, 1
FRONT_HALO, BACK_HALO = 8, 136


def _ctype(decl):
    decl = decl.strip()
    if "*" in decl:
        return ctypes.c_void_p
    base = re.sub(r"\b(const|unsigned)\b", "", decl).split()
    ty = base[0]
    return {"int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "double": ctypes.c_double,
            "size_t": ctypes.c_size_t, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32}[ty]


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes], [argnames])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int|size_t|long)\s+(dcue_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else {"size_t": ctypes.c_size_t, "long": ctypes.c_long}.get(ret, ctypes.c_int)
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                argtypes.append(_ctype(a))
                argnames.append(re.findall(r"\w+", a)[-1])
        protos[name] = (restype, argtypes, argnames)
    return protos


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libdcue_b200.so not found at %s: build it with `python __graft_entry__.py` "
                "(there is no CPU / eager fallback for the DCUE hot path)" % LIB_PATH)
        _protos = parse_header()
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args, _) in _protos.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.argtypes = args
            fn.restype = res
        _lib = handle
    return _lib


class DcueError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().dcue_last_error()
        msg = msg.decode() if msg else ""
        raise DcueError("%s failed with code %d: %s" % (what or "dcue kernel", rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    fn = getattr(lib(), name)
    if len(args) != len(fn.argtypes):
        raise TypeError("%s expects %d arguments, got %d" % (name, len(fn.argtypes), len(args)))
    check(fn(*args), name)


def query(name, *args):
    """For the size_t-returning workspace queries."""
    return int(getattr(lib(), name)(*args))
