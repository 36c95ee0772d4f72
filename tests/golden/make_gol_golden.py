"""Generates tests/golden/gol_golden.npz by running the REFERENCE Game-of-Life worlds
(/root/reference/game_of_life/envs/world_pytorch.py and world.py, imported by path) in this
container.  Run once: python tests/golden/make_gol_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/game_of_life/envs"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    wp, wo = load("world_pytorch"), load("world")
    out = {}
    # --- multi world: N boards, toggle + tick per step (multi_env.py:170-232 around World._tick)
    for (n, w, steps, seed) in [(8, 16, 64, 0), (4, 5, 40, 1), (3, 32, 24, 2), (5, 8, 50, 3)]:
        rng = np.random.default_rng(seed)
        init = (rng.random((n, w, w)) < 0.2).astype(np.uint8)
        acts = rng.integers(0, w * w, size=(steps, n))
        world = wp.World(map_width=w, map_height=w, prob_life=20, cuda=False, num_proc=n)
        world.state = torch.from_numpy(init.astype(np.float32)).reshape(n, 1, w, w)
        states, pops = [], []
        for t in range(steps):
            onehot = torch.zeros(n, w * w, dtype=torch.long)
            onehot[torch.arange(n), torch.from_numpy(acts[t])] = 1
            world.state = (world.state.long() ^ onehot.view(n, 1, w, w)).float()
            world._tick()
            s = world.state.reshape(n, w, w).numpy().astype(np.uint8)
            states.append(s)
            pops.append(s.reshape(n, -1).sum(1))
        k = "multi_%d_%d" % (n, w)
        out[k + "_init"] = init
        out[k + "_acts"] = acts
        out[k + "_states"] = np.stack(states)
        out[k + "_pops"] = np.stack(pops)
    # --- classic object world with stale neighbour references (world.py)
    for (w, steps, seed) in [(16, 80, 0), (6, 60, 1)]:
        rng = np.random.default_rng(100 + seed)
        init = (rng.random((w, w)) < 0.21).astype(np.uint8)
        acts = rng.integers(0, w * w, size=steps)
        np.random.seed(seed)
        world = wo.World(w, w, prob_life=20)
        for x in range(w):
            for y in range(w):
                world.build_cell(x, y, alive=bool(init[x, y]))
        world.prepopulate_neighbours()
        states, raws = [], []
        for t in range(steps):
            a = int(acts[t])
            raws.append(int(world.state.sum()))
            world.build_cell(a // w, a % w, alive=True)
            world._tick()
            states.append(world.state[0].copy())
        k = "classic_%d" % w
        out[k + "_init"] = init
        out[k + "_acts"] = acts
        out[k + "_states"] = np.stack(states)
        out[k + "_raw"] = np.array(raws)
    np.savez_compressed(os.path.join(HERE, "gol_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "gol_golden.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
