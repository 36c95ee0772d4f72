"""gymcity_b200 -- B200-native batched environment tick for gym-city (Micropolis + Game of Life).

Only the tick path: CUDA kernels + C-ABI in csrc/, and host-side mirrors of the reference's
MicropolisEnv / MicropolisControl / GoL env interfaces.  No CPU fallback.
"""
from ._lib import lib, NativeLibraryMissing, LIB_PATH  # noqa: F401


def __getattr__(name):
    if name in ("GoLMultiEnv", "GameOfLifeEnv"):
        from . import gol_env
        return getattr(gol_env, name)
    if name in ("MicropolisEnv", "MicropolisControl", "MicropolisVecEnv", "MicropolisPaintEnv"):
        from . import micropolis_env
        return getattr(micropolis_env, name)
    if name in ("make_vec_envs", "RolloutStorage", "BatchedVecEnv"):
        from . import envs
        return getattr(envs, name)
    if name == "Extinguisher":
        from . import wrappers
        return wrappers.Extinguisher
    raise AttributeError(name)
