"""DQN / DDQN data path on the device: replay ring with bit-packed windows and the masked
epsilon-greedy action selection of agents/ddqn_agent.py:95-108 for a whole batch of envs.

    env = MazeVectorEnv(B, shape, enrich=True, ...)
    memory = DeviceReplay(env, capacity)         # lib/replay_memory.py:8-24 on the device
    actor = MaskedEpsilonGreedy(env, starting_epsilon, final_epsilon, epsilon_decay)
    env.reset(); memory.observe()
    for step in ...:
        q = policy_net(memory.current_state())   # any torch module: (vec [B,6], window [B,3,15,15]) -> [B,4]
        a = actor.select(q)
        env.step(a); memory.push(a)
        state, action, reward, next_state = memory.sample(batch_size)

The neural network itself is the consumer's (PyTorch); everything either side of it runs in the
kernels of csrc/maze_dqn.cu.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import cabi
from .agents import epsilon_lut


class DeviceReplay:
    def __init__(self, env, capacity: int, seed: int = 0, without_replacement: bool = True):
        """without_replacement=True: a sampled batch holds distinct transitions, like random.sample in the reference's
        ReplayMemory.sample (lib/replay_memory.py:20-21); False: independent uniform draws."""
        self.batch = env.batch if hasattr(env, "batch") else env
        self.device, self.ctx = self.batch.device, self.batch.ctx
        self.capacity = int(capacity)
        if self.capacity < self.batch.num_envs:
            raise ValueError("capacity must be at least num_envs (one push launch claims up to num_envs slots of the ring)")
        B, d, W = self.batch.num_envs, self.device, cabi.WINDOW_WORDS
        self.pushed = torch.zeros(1, dtype=torch.int64, device=d)
        self.vec = torch.zeros((self.capacity, 6), dtype=torch.float32, device=d)
        self.next_vec = torch.zeros((self.capacity, 6), dtype=torch.float32, device=d)
        self.win = torch.zeros((self.capacity, W), dtype=torch.int32, device=d)
        self.next_win = torch.zeros((self.capacity, W), dtype=torch.int32, device=d)
        self.action = torch.zeros(self.capacity, dtype=torch.uint8, device=d)
        self.reward = torch.zeros(self.capacity, dtype=torch.float32, device=d)
        self.stage_vec = torch.zeros((B, 6), dtype=torch.float32, device=d)
        self.stage_win = torch.zeros((B, W), dtype=torch.int32, device=d)
        self.seed, self._draw = int(seed), 0
        self._c = cabi.MazeReplay(
            capacity=self.capacity, pushed=self.pushed.data_ptr(), vec=self.vec.data_ptr(), next_vec=self.next_vec.data_ptr(),
            win=self.win.data_ptr(), next_win=self.next_win.data_ptr(), action=self.action.data_ptr(),
            reward=self.reward.data_ptr(), stage_vec=self.stage_vec.data_ptr(), stage_win=self.stage_win.data_ptr(),
            without_replacement=int(bool(without_replacement)), reserved=0)

    def _stream(self):
        return cabi.current_stream(self.device)

    def __len__(self):
        return min(int(self.pushed.item()), self.capacity)

    def observe(self):
        """Stage the current observation of every env (call after env.reset())."""
        self.batch.sync_flags()
        rc = cabi.lib().maze_dqn_observe(self.ctx.handle, C.byref(self.batch._c), C.byref(self._c), self._stream())
        self.ctx.check(rc, "maze_dqn_observe")

    def push(self, actions: torch.Tensor):
        """memorize(state, action, reward, next_state) for the step just made with `actions`."""
        if actions.dtype != torch.uint8 or actions.device != self.device or not actions.is_contiguous():
            actions = actions.to(device=self.device, dtype=torch.uint8).contiguous()
        self.batch.sync_flags()
        rc = cabi.lib().maze_dqn_push(self.ctx.handle, C.byref(self.batch._c), C.byref(self._c), cabi.ptr(actions), self._stream())
        self.ctx.check(rc, "maze_dqn_push")

    def current_state(self):
        """(vec [B, 6] float32, window [B, 3, 15, 15] float32) of the staged observation."""
        return self.stage_vec, unpack_windows(self.stage_win)

    def sample_packed(self, n: int, check: bool = False):
        """memory.sample(n) for the tensor-core net (maze_b200.dqn_net.DQNNet): windows stay bit-packed.
        -> vec [n, 6] f32, win [n, 24] i32, next_vec, next_win, action [n] u8, reward [n] f32 (the same transitions
        sample() returns for the same draw).  check=True raises when fewer than n transitions are stored (one sync)."""
        if check and len(self) < n:
            raise ValueError(f"replay holds {len(self)} transitions, fewer than the batch of {n} (ddqn_agent.py:114-115 returns)")
        d, W = self.device, cabi.WINDOW_WORDS
        if getattr(self, "_packed_out", None) is None or self._packed_out["vec"].shape[0] != n:
            self._packed_out = dict(vec=torch.zeros((n, 6), dtype=torch.float32, device=d), win=torch.zeros((n, W), dtype=torch.int32, device=d),
                                    next_vec=torch.zeros((n, 6), dtype=torch.float32, device=d), next_win=torch.zeros((n, W), dtype=torch.int32, device=d),
                                    action=torch.zeros(n, dtype=torch.uint8, device=d), reward=torch.zeros(n, dtype=torch.float32, device=d))
        o = self._packed_out
        self._draw += 1
        rc = cabi.lib().maze_dqn_sample_packed(self.ctx.handle, C.byref(self._c), int(n), self.seed & (2**64 - 1), self._draw,
                                               cabi.ptr(o["vec"]), cabi.ptr(o["win"]), cabi.ptr(o["next_vec"]), cabi.ptr(o["next_win"]),
                                               cabi.ptr(o["action"]), cabi.ptr(o["reward"]), self._stream())
        self.ctx.check(rc, "maze_dqn_sample_packed")
        return o["vec"], o["win"], o["next_vec"], o["next_win"], o["action"], o["reward"]

    def sample(self, n: int, check: bool = True):
        """-> (vec, window), action [n] int64, reward [n] float32, (next_vec, next_window).  Raises when the ring is
        empty (check=True costs one device sync; the reference returns early while len(memory) < batch_size,
        ddqn_agent.py:114-115)."""
        if check and len(self) == 0:
            raise ValueError("the replay ring is empty")
        d = self.device
        out = dict(vec=torch.zeros((n, 6), dtype=torch.float32, device=d), win=torch.zeros((n, 3, cabi.WINDOW, cabi.WINDOW), dtype=torch.float32, device=d),
                   next_vec=torch.empty((n, 6), dtype=torch.float32, device=d),
                   next_win=torch.empty((n, 3, cabi.WINDOW, cabi.WINDOW), dtype=torch.float32, device=d),
                   action=torch.empty(n, dtype=torch.int64, device=d), reward=torch.empty(n, dtype=torch.float32, device=d))
        self._draw += 1
        rc = cabi.lib().maze_dqn_sample(self.ctx.handle, C.byref(self._c), int(n), self.seed & (2**64 - 1), self._draw,
                                        cabi.ptr(out["vec"]), cabi.ptr(out["win"]), cabi.ptr(out["next_vec"]), cabi.ptr(out["next_win"]),
                                        cabi.ptr(out["action"]), cabi.ptr(out["reward"]), self._stream())
        self.ctx.check(rc, "maze_dqn_sample")
        return (out["vec"], out["win"]), out["action"], out["reward"], (out["next_vec"], out["next_win"])


def unpack_windows(words: torch.Tensor) -> torch.Tensor:
    """[n, 24] packed windows -> [n, 3, 15, 15] float32 (host-side helper built from torch ops; the
    sampling kernel unpacks on its own)."""
    n = words.shape[0]
    w = words.view(n, 3, 8, 1).to(torch.int64) & 0xffffffff
    bits = (w >> torch.arange(32, device=words.device).view(1, 1, 1, 32)) & 1
    # word k of a channel = window rows 2 k (bits 0-14) and 2 k + 1 (bits 16-30)
    rows = bits.reshape(n, 3, 16, 16)[:, :, :cabi.WINDOW, :cabi.WINDOW]
    return rows.to(torch.float32).contiguous()


class MaskedEpsilonGreedy:
    """DDQNAgent.get_action (ddqn_agent.py:98-108) for every env: epsilon from each env's own
    steps_done; exploration draws from get_mask_direction(probs=True) / sum."""

    def __init__(self, env, starting_epsilon: float, final_epsilon: float, epsilon_decay: float, seed: int = 0, env_id_base: int = 0):
        self.batch = env.batch if hasattr(env, "batch") else env
        self.device, self.ctx = self.batch.device, self.batch.ctx
        self.eps_lut = torch.from_numpy(epsilon_lut(starting_epsilon, final_epsilon, epsilon_decay)).to(self.device)
        self.steps_done = torch.zeros(self.batch.num_envs, dtype=torch.int32, device=self.device)
        self.actions = torch.zeros(self.batch.num_envs, dtype=torch.uint8, device=self.device)
        self.seed, self.env_id_base = int(seed), int(env_id_base)

    def select(self, q_values: torch.Tensor) -> torch.Tensor:
        q = q_values.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if q.shape != (self.batch.num_envs, 4):
            raise ValueError(f"q_values must be [{self.batch.num_envs}, 4]")
        rc = cabi.lib().maze_dqn_select(self.ctx.handle, C.byref(self.batch._c), cabi.ptr(q), cabi.ptr(self.eps_lut), self.eps_lut.numel(),
                                        cabi.ptr(self.steps_done), self.seed & (2**64 - 1), self.env_id_base,
                                        cabi.ptr(self.actions), cabi.current_stream(self.device))
        self.ctx.check(rc, "maze_dqn_select")
        return self.actions
