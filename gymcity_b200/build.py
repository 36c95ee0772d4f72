"""Builds gymcity_b200/libgymcity_b200.so (sm_100a) in-tree with nvcc.

`python -m gymcity_b200.build` or `__graft_entry__.build()`.  The library is the product: there is
no CPU fallback, and `gymcity_b200._lib` raises if it is missing.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("GYMCITY_B200_LIB", os.path.join(HERE, "libgymcity_b200.so"))
SOURCES = ["gol.cu", "dlpack_export.cu", "mp_batch.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exact float: no fused multiply-add contraction, IEEE div/sqrt (setValves, collectTax,
    # getScore in the reference are plain x86-64 SSE2 scalar code)
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MPB_NVCC_EXTRA", "").split()      # e.g. -DMP_PROFILE_SPLIT for profiling builds
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h", ".cuh"))]    # sources, not objects
    inc = os.path.join(HERE, "..", "include")
    deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    if not force and not _newer(LIB, deps):
        return LIB
    objs = []
    procs = []
    for s in srcs:
        o = s[:-3] + os.environ.get("MPB_OBJ_SUFFIX", "") + ".o"
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + extra + ["-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((subprocess.Popen(cmd), cmd))
        objs.append(o)
    for p, cmd in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
           "-o", LIB] + objs
    subprocess.check_call(cmd)
    return LIB


def build_coverage():
    """The -DMP_COVERAGE variant (csrc/mp_cov.h) next to the product library: test infrastructure for
    tests/test_coverage_gpu.py, never loaded unless GYMCITY_B200_LIB names it."""
    global LIB
    keep, env_keep = LIB, {k: os.environ.get(k) for k in ("MPB_NVCC_EXTRA", "MPB_OBJ_SUFFIX")}
    LIB = os.path.join(HERE, "libgymcity_b200_cov.so")
    os.environ["MPB_NVCC_EXTRA"] = "-DMP_COVERAGE"
    os.environ["MPB_OBJ_SUFFIX"] = "_cov"
    try:
        return build()
    finally:
        LIB = keep
        for k, v in env_keep.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--coverage" in sys.argv:
        print(build_coverage())
