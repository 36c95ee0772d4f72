"""Zero-copy torch views of device buffers owned by a native handle (DLPack capsules)."""
import ctypes as C

import numpy as np

from ._lib import lib

_CODES = {"int8": (0, 8), "int16": (0, 16), "int32": (0, 32), "int64": (0, 64),
          "uint8": (1, 8), "uint16": (1, 16), "uint32": (1, 32), "float32": (2, 32)}

C.pythonapi.PyCapsule_New.restype = C.py_object
C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]


def device_tensor(ptr, dtype, shape, device):
    """torch tensor aliasing `ptr` (device memory) -- no copy.  The owner must outlive it."""
    import torch
    code, bits = _CODES[dtype]
    shp = (C.c_int64 * len(shape))(*[int(s) for s in shape])
    L = lib()
    L.gcb_dlpack_wrap.restype = C.c_void_p
    L.gcb_dlpack_wrap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    mt = L.gcb_dlpack_wrap(ptr, code, bits, len(shape), shp, int(device))
    if not mt:
        raise RuntimeError("gcb_dlpack_wrap failed")
    cap = C.pythonapi.PyCapsule_New(mt, b"dltensor", None)
    return torch.from_dlpack(cap)
