"""ctypes binding of libsgb200.so (the C-ABI declared in include/sgb200.h).

The signature table is parsed from the header, so the binding and the header cannot drift apart.  There is NO
fallback: if the library is missing on a machine with a GPU the import fails loudly (`python -m csrc.build` or
`__graft_entry__.build()` produces it)."""
import ctypes
import os
import re

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "lib", "libsgb200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "sgb200.h")

F32, BF16 = 0, 1
DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16}


class SgConvGeom(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int32), ("Hq", ctypes.c_int32), ("Wq", ctypes.c_int32),
                ("Hin", ctypes.c_int32), ("Win", ctypes.c_int32), ("Cin", ctypes.c_int32),
                ("Hout", ctypes.c_int32), ("Wout", ctypes.c_int32), ("Cout", ctypes.c_int32),
                ("in_stride", ctypes.c_int32), ("out_stride", ctypes.c_int32),
                ("nphase", ctypes.c_int32), ("ntaps", ctypes.c_int32),
                ("dy", (ctypes.c_int8 * 16) * 4), ("dx", (ctypes.c_int8 * 16) * 4)]


_CTYPE = {"int": ctypes.c_int, "float": ctypes.c_float, "long long": ctypes.c_longlong,
          "unsigned long long": ctypes.c_ulonglong, "double": ctypes.c_double}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?m)^\s*(const char\*|long long|int)\s+(sg_\w+)\s*\(([^;]*?)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0].replace("const ", "").strip()
                    argtypes.append(_CTYPE[ty])
        out[name] = (ctypes.c_char_p if ret.startswith("const char") else
                     (ctypes.c_longlong if ret == "long long" else ctypes.c_int), argtypes)
    return out


class _Lib:
    def __init__(self):
        self._dll = None
        self.signatures = parse_header()

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "sgb200: %s is missing -- build it with `python stylegan.pytorch_b200/csrc/build.py` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        dll = ctypes.CDLL(LIB_PATH)
        for name, (ret, argtypes) in self.signatures.items():
            fn = getattr(dll, name)     # AttributeError here = header/library drift
            fn.restype = ret
            fn.argtypes = argtypes
        if dll.sg_abi_version() != 2 or dll.sg_geom_size() != ctypes.sizeof(SgConvGeom):
            raise RuntimeError("sgb200: ABI mismatch between models/_lib.py and libsgb200.so")
        self._dll = dll
        return dll

    def __getattr__(self, name):
        if name.startswith("sg_"):
            dll = self.load()
            raw = getattr(dll, name)
            if self.signatures[name][0] is ctypes.c_int and name not in _PLAIN_INT:
                def call(*args, _raw=raw, _name=name):
                    rc = _raw(*args)
                    if rc != 0:
                        raise RuntimeError("sgb200 %s failed (%d): %s" % (_name, rc, dll.sg_last_error().decode()))
                fn = call
            else:
                fn = raw
            setattr(self, name, fn)
            return fn
        raise AttributeError(name)


# functions whose int return value is a result, not a status
_PLAIN_INT = {"sg_lrelu_bwd_bias_supported", "sg_fast_paths", "sg_pair_updown", "sg_abi_version", "sg_geom_size", "sg_entry_size", "sg_wgrad_splits", "sg_channel_sum_blocks",
              "sg_chunk_elems", "sg_conv3x3_tc_supported", "sg_conv3x3_pair_supported", "sg_conv3x3_pair_slots", "sg_blur3_epi_slots", "sg_lingroup_rows", "sg_lingroup_supported", "sg_lingroup_fwd_blocks",
              "sg_lingroup_wgrad_blocks", "sg_wgrad3x3_tc_supported", "sg_wgrad3x3_tc_splits", "sg_wgrad3x3_v2_supported", "sg_wgrad3x3_v2_splits", "sg_convud_tc_supported", "sg_wgradud_tc_supported", "sg_wgradud_tc_splits"}

lib = _Lib()

_launches = 0


def count_launch(n=1):
    global _launches
    _launches += n


def launches():
    return _launches


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def code(dtype):
    return DTYPE_CODE[dtype]
