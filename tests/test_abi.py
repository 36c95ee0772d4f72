"""CPU: the C-ABI library loads and exports every symbol include/gymcity_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gymcity_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:golb|mpb|gcb)_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from gymcity_b200 import LIB_PATH
    L = ctypes.CDLL(LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    from gymcity_b200 import lib
    assert not lib().golb_create(4, 16, 10, 0, 0)      # NULL, no silent CPU path
