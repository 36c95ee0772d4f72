"""CPU harness (TEST INFRASTRUCTURE): builds tests/hostbuild/host_engine.cpp -- the same engine
headers the CUDA kernels compile, built with g++ for one host thread -- into
tests/_hostbuild/libmph.so and wraps it with the oracle's EngineBase ABI so the CPU suite can diff
the engine logic against the reference engine.  gymcity_b200 itself never loads this."""
import ctypes as C
import glob
import os
import subprocess

from oracle.ref_engine import EngineBase

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "hostbuild", "host_engine.cpp")
LIB = os.path.join(HERE, "_hostbuild", "libmph.so")


def build_host_engine():
    deps = [SRC] + glob.glob(os.path.join(ROOT, "gymcity_b200", "csrc", "mp_*.h")) \
        + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", LIB, SRC])
    return LIB


class HostEngine(EngineBase):
    def __init__(self):
        super().__init__(C.CDLL(build_host_engine()), "mph_")
