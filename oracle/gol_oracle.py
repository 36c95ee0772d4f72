"""CPU restatement of the reference Game-of-Life step -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  Pinned against the reference itself: tests/golden/gol_golden.npz was produced by
tests/golden/make_gol_golden.py importing /root/reference/game_of_life/envs/world_pytorch.py and
world.py by path (tests/test_gol_oracle.py checks this module against those vectors).
"""
import numpy as np


def tick_torus(state):
    """world_pytorch.py:91-135: circular pad, conv [[1,1,1],[1,9,1],[1,1,1]], GoLU.
    state: (..., W, H) integer array of 0/1.  v = n + 9*self; alive iff v in {3, 11, 12}."""
    s = np.asarray(state).astype(np.int32)
    n = np.zeros_like(s)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            if dx or dy:
                n += np.roll(np.roll(s, dx, axis=-2), dy, axis=-1)
    v = n + 9 * s
    return ((v == 3) | (v == 11) | (v == 12)).astype(np.uint8)


def multi_step(state, actions, trg_pop, max_loss, max_step, num_proc):
    """multi_env.py:170-232 for state (N, W, W) uint8 and actions (N,) flat indices.
    Returns (new_state, reward float32 (N,), pop int (N,))."""
    n, w, _ = state.shape
    st = state.copy()
    a = np.asarray(actions).reshape(-1)
    st[np.arange(n), a // w, a % w] ^= 1                   # :179,243-253
    st = tick_torus(st)                                     # :193-194 (sim_steps = 1)
    pop = st.reshape(n, -1).sum(1)
    f = np.float32
    loss = np.abs(f(trg_pop) - pop.astype(np.float32))      # :203-204
    denom = f(f(f(max_loss) * f(max_step)) * f(num_proc))
    reward = ((f(max_loss) - loss) * f(100)) / denom        # :208
    return st, reward.astype(np.float32), pop


class ClassicCell:
    __slots__ = ("x", "y", "alive", "next_state", "neighbours")

    def __init__(self, x, y, alive):
        self.x, self.y, self.alive, self.next_state, self.neighbours = x, y, alive, None, None


class ClassicWorld:
    """world.py:3-158 semantics: per-cell objects with cached neighbour references that go stale
    when build_cell replaces a cell (world.py:108-112,119-130)."""
    DIRS = [(-1, 1), (0, 1), (1, 1), (-1, 0), (1, 0), (-1, -1), (0, -1), (1, -1)]  # world.py:16-20

    def __init__(self, board):
        b = np.asarray(board)
        self.w = b.shape[0]
        self.cells = [[ClassicCell(x, y, bool(b[x, y])) for y in range(self.w)]
                      for x in range(self.w)]
        for row in self.cells:                               # prepopulate_neighbours
            for c in row:
                self._neigh(c)

    def _neigh(self, c):
        if c.neighbours is None:
            c.neighbours = [self.cells[(c.x + dx) % self.w][(c.y + dy) % self.w]
                            for dx, dy in self.DIRS]
        return c.neighbours

    def build_cell(self, x, y, alive=True):
        self.cells[x][y] = ClassicCell(x, y, alive)

    def state(self):
        return np.array([[int(c.alive) for c in row] for row in self.cells], dtype=np.uint8)

    def tick(self):
        for row in self.cells:
            for c in row:
                n = sum(1 for q in self._neigh(c) if q.alive)
                if c.alive is False and n == 3:
                    c.next_state = 1
                elif n < 2 or n > 3:
                    c.next_state = 0
        for row in self.cells:
            for c in row:
                if c.next_state == 1:
                    c.alive = True
                elif c.next_state == 0:
                    c.alive = False


def classic_step(world, a, max_step, step_count):
    """env.py:125-182 (non-prebuild path).  Returns (state, reward, terminal)."""
    w = world.w
    dele = a < 0
    a = abs(a)
    x, y = a // w, a % w
    if dele:
        world.build_cell(x, y, alive=False)
    raw = int(world.state().sum())
    if not dele:
        world.build_cell(x, y, alive=True)
    world.tick()
    terminal = (step_count == max_step) or raw < 2
    return world.state(), float(raw * 100 / (max_step * w * w)), terminal


class TorchWorldPort:
    """world_pytorch.py:6-135 + multi_env.py:170-232 restated with the reference's own torch ops on the CPU (all of
    torch's threads): circular pad by four torch.cat, Conv2d(1, 1, 3) with weights [[1,1,1],[1,9,1],[1,1,1]], round,
    the piecewise GoLU activation, population by sum, fp32 reward.  bench.py's CPU baseline for the Game-of-Life
    workload (the reference's own CPU path is exactly this module with cuda=False).  Checked against tick_torus
    in tests/test_gol_oracle.py."""

    def __init__(self, state):
        import torch
        self.torch = torch
        self.state = torch.as_tensor(np.asarray(state), dtype=torch.float32).unsqueeze(1).contiguous()   # (N, 1, W, W)
        self.weight = torch.tensor([[[[1, 1, 1], [1, 9, 1], [1, 1, 1]]]], dtype=torch.float32)
        n, _, w, _ = self.state.shape
        self.n, self.w = n, w

    @staticmethod
    def _pad_circular(torch, x, pad):                                  # world_pytorch.py:6-19
        x = torch.cat([x, x[:, :, 0:pad, :]], dim=2)
        x = torch.cat([x, x[:, :, :, 0:pad]], dim=3)
        x = torch.cat([x[:, :, -2 * pad:-pad, :], x], dim=2)
        x = torch.cat([x[:, :, :, -2 * pad:-pad], x], dim=3)
        return x

    def _golu(self, x):                                                # world_pytorch.py:110-135
        x_out = x.clone().fill_(0).float()
        ded_0 = (x >= 2).float()
        bth_0 = ded_0 * (x < 3).float()
        x_out = x_out + (bth_0 * (x - 2).float())
        ded_1 = (x >= 3).float()
        bth_1 = ded_1 * (x < 4).float()
        x_out = x_out + abs(bth_1 * (x - 4).float())
        alv_0 = (x >= 10).float()
        lif_0 = alv_0 * (x < 11).float()
        x_out = x_out + (lif_0 * (x - 10).float())
        alv_1 = (x >= 11).float()
        lif_1 = alv_1 * (x < 12).float()
        x_out = x_out + lif_1
        alv_2 = (x >= 12).float()
        lif_2 = alv_2 * (x < 13).float()
        x_out = x_out + abs(lif_2 * (x - 13).float())
        return x_out

    def step(self, actions, trg_pop, max_loss, max_step):
        torch = self.torch
        with torch.no_grad():
            a = torch.as_tensor(np.asarray(actions), dtype=torch.int64)
            onehot = torch.zeros(self.n, self.w * self.w)
            onehot.scatter_(1, a.reshape(-1, 1), 1.0)                   # multi_env.py:243-253 act_to_build
            st = ((self.state.long() ^ onehot.reshape(self.n, 1, self.w, self.w).long())).float()   # :179
            x = self._pad_circular(torch, st, 1)
            x = torch.nn.functional.conv2d(x, self.weight)
            x = x.round()
            self.state = self._golu(x)
            pop = self.state.sum(dim=(1, 2, 3))                         # :199-204
            loss = abs(trg_pop - pop)
            reward = (max_loss - loss) * 100 / (max_loss * max_step * self.n)
        return self.state, reward, pop
