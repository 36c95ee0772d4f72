"""ctypes binding of libgymcity_b200.so (the C-ABI in include/gymcity_b200.h).

No fallback: if the CUDA library is missing or cannot be loaded this raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GYMCITY_B200_LIB", os.path.join(HERE, "libgymcity_b200.so"))   # override: perf variants

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "%s not built: run `python -m gymcity_b200.build` (needs nvcc; sm_100a only). "
            "gymcity_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, f32 = C.c_void_p, C.c_int, C.c_uint32, C.c_float
    sig = {
        "golb_create": (vp, [i32, i32, i32, i32, i32]),
        "golb_destroy": (None, [vp]),
        "golb_last_error": (C.c_char_p, [vp]),
        "golb_set_target": (i32, [vp, f32, f32]),
        "golb_reset": (i32, [vp, u32, f32, vp]),
        "golb_set_state": (i32, [vp, vp, vp]),
        "golb_get_state": (i32, [vp, vp]),
        "golb_step": (i32, [vp, vp, vp]),
        "golb_step_host": (i32, [vp, vp, vp, vp]),
        "golb_step_n": (i32, [vp, vp, i32, vp, vp, vp]),
        "golb_ptr": (vp, [vp, i32]),
        "golb_step_count": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
