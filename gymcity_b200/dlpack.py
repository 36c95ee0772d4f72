"""Zero-copy torch views of device buffers owned by a native handle (DLPack capsules).

Lifetime: a view keeps its owner alive.  `device_tensor(..., owner=obj)` registers the view with `obj` (a `ViewOwner`);
the DLPack deleter -- run when torch frees the tensor's storage, i.e. when the last tensor sharing it is gone --
unregisters it.  While views exist the owner object is referenced from here (so it is not garbage-collected under
them), and an explicit `close()` is deferred until the last view is released: no tensor handed out by an env can
outlive the device memory it aliases."""
import ctypes as C
import itertools
import threading

from ._lib import lib

_CODES = {"int8": (0, 8), "int16": (0, 16), "int32": (0, 32), "int64": (0, 64),
          "uint8": (1, 8), "uint16": (1, 16), "uint32": (1, 32), "float32": (2, 32)}

C.pythonapi.PyCapsule_New.restype = C.py_object
C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]

_RELEASE = C.CFUNCTYPE(None, C.c_void_p)
_live = {}                       # view id -> owner
_ids = itertools.count(1)        # next() is atomic: envs may be created from several threads
_lock = threading.RLock()        # guards the view counts (a view may be released on another thread than its owner's)


class ViewOwner:
    """Mixin for objects that own a native handle: `_destroy_native()` frees it; `close()` does so at once when no view
    of its memory is alive, otherwise when the last one is released."""
    _views = 0
    _close_pending = False

    def _destroy_native(self):
        raise NotImplementedError

    def close(self):
        with _lock:
            if self._views > 0:
                self._close_pending = True
                return
        self._destroy_native()

    def _view_released(self):
        with _lock:
            self._views -= 1
            fire = self._views <= 0 and self._close_pending
            if fire:
                self._close_pending = False
        if fire:
            self._destroy_native()


@_RELEASE
def _on_release(ctx):
    try:
        owner = _live.pop(int(ctx or 0), None)
        if owner is not None:
            owner._view_released()
    except Exception:            # interpreter shutdown: nothing left to protect
        pass


def device_tensor(ptr, dtype, shape, device, owner=None):
    """torch tensor aliasing `ptr` (device memory) -- no copy.  With `owner` (a ViewOwner) the view keeps it alive."""
    import torch
    code, bits = _CODES[dtype]
    shp = (C.c_int64 * len(shape))(*[int(s) for s in shape])
    L = lib()
    L.gcb_dlpack_wrap.restype = C.c_void_p
    L.gcb_dlpack_wrap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.gcb_dlpack_wrap_owned.restype = C.c_void_p
    L.gcb_dlpack_wrap_owned.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, _RELEASE, C.c_void_p]
    if owner is None:
        mt = L.gcb_dlpack_wrap(ptr, code, bits, len(shape), shp, int(device))
    else:
        vid = next(_ids)
        with _lock:
            _live[vid] = owner
            owner._views += 1
        mt = L.gcb_dlpack_wrap_owned(ptr, code, bits, len(shape), shp, int(device), _on_release, C.c_void_p(vid))
        if not mt:
            with _lock:
                _live.pop(vid, None)
                owner._views -= 1
    if not mt:
        raise RuntimeError("gcb_dlpack_wrap failed")
    cap = C.pythonapi.PyCapsule_New(mt, b"dltensor", None)
    return torch.from_dlpack(cap)
