"""Tabular agents on the device: the reference's QAgent / DQAgent (agents/q_agent.py:8-79,
agents/dq_agent.py:5-73) over a batch of environments.

`TabularAgent` owns the hash-table Q function in HBM and drives the kernels behind
maze_q_act / maze_q_update / maze_q_rollout.  `QAgent` and `DQAgent` keep the reference's
constructor signature (env, learning_rate, initial_epsilon, epsilon_decay, final_epsilon,
discount_factor, eta) and method names (get_action, update, update_hyperparameter) with batched
meaning: `env` is a MazeVectorEnv (or anything with a `.batch`), observations are implicit (the
learner reads the env state on the device), actions are uint8 [B] device tensors.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import cabi


def epsilon_lut(initial_epsilon: float, final_epsilon: float, decay: float, max_len: int = 1 << 22) -> np.ndarray:
    """eps[steps_done] exactly as q_agent.py:49 evaluates it, up to the point where it has
    converged to final_epsilon in float64 (or max_len)."""
    n = int(min(max_len, max(2, math.ceil(40.0 * decay) + 2)))
    out = (C.c_double * n)()
    rc = cabi.lib().maze_q_epsilon_lut(float(initial_epsilon), float(final_epsilon), float(decay), out, n)
    if rc != 0:
        raise cabi.MazeError(f"maze_q_epsilon_lut -> {rc}")
    return np.array(out, dtype=np.float64)


class TabularAgent:
    def __init__(self, batch, learning_rate: float, initial_epsilon: float, epsilon_decay: float, final_epsilon: float,
                 discount_factor: float, eta: float, double_q: bool = False, envs_per_agent: int = 1,
                 capacity: int | None = None, seed: int = 0, env_id_base: int = 0):
        self.batch = batch
        self.device = batch.device
        self.ctx = batch.ctx
        B = batch.num_envs
        self.envs_per_agent = int(envs_per_agent)
        if not 1 <= self.envs_per_agent <= B:
            raise ValueError("envs_per_agent must be in [1, num_envs]")
        self.num_agents = (B + self.envs_per_agent - 1) // self.envs_per_agent
        if capacity is None:
            # every (agent, open block, best-dir) can become a row; keep the load factor under 1/2
            per_agent = batch.pool.slot // 2 + 16
            capacity = 2 * per_agent * max(self.num_agents, min(batch.pool.num_mazes, B))
        cap = 16
        while cap < capacity:
            cap <<= 1
        if cap > 1 << 31:
            raise ValueError("Q table capacity above 2^31 rows")
        self.capacity = cap
        d = self.device
        self.keys = torch.full((cap,), -1, dtype=torch.int64, device=d)          # MAZE_Q_EMPTY
        self.q_a = torch.zeros((cap, 4), dtype=torch.float64, device=d)
        self.q_b = torch.zeros((cap, 4), dtype=torch.float64, device=d) if double_q else None
        self.overflow = torch.zeros(1, dtype=torch.int32, device=d)
        self.eps_host = epsilon_lut(initial_epsilon, final_epsilon, epsilon_decay)
        self.eps_lut = torch.from_numpy(self.eps_host).to(d)
        self.gamma = torch.full((self.num_agents,), float(discount_factor), dtype=torch.float64, device=d)
        self.lr, self.eta = float(learning_rate), float(eta)
        self.slot = torch.full((B,), -1, dtype=torch.int32, device=d)             # MAZE_Q_NO_SLOT
        self.steps_done = torch.zeros(B, dtype=torch.int32, device=d)
        self.last_action = torch.zeros(B, dtype=torch.uint8, device=d)
        self.ep_return = torch.zeros(B, dtype=torch.float64, device=d)
        self.actions = torch.zeros(B, dtype=torch.uint8, device=d)
        self.seed, self.env_id_base = int(seed), int(env_id_base)
        self._tapes = None
        self._c = self._make_struct()

    def _make_struct(self):
        t = self._tapes
        return cabi.MazeQAgent(
            capacity=self.capacity, keys=self.keys.data_ptr(), q_a=self.q_a.data_ptr(),
            q_b=None if self.q_b is None else self.q_b.data_ptr(), overflow=self.overflow.data_ptr(),
            envs_per_agent=self.envs_per_agent, eps_len=self.eps_lut.numel(), eps_lut=self.eps_lut.data_ptr(),
            gamma=self.gamma.data_ptr(), lr=self.lr, eta=self.eta, slot=self.slot.data_ptr(),
            steps_done=self.steps_done.data_ptr(), last_action=self.last_action.data_ptr(),
            ep_return=self.ep_return.data_ptr(), seed=self.seed & (2**64 - 1), env_id_base=self.env_id_base,
            u_tape=None if t is None else t[0].data_ptr(), a_tape=None if t is None else t[1].data_ptr(),
            tape_pos=None if t is None else t[2].data_ptr(), u_len=0 if t is None else t[0].shape[0],
            a_len=0 if t is None else t[1].shape[0])

    def attach_tapes(self, u_tape: np.ndarray, a_tape: np.ndarray):
        """Replay recorded draws instead of Philox (tests): u_tape [Tu, B] float64 = the values
        np.random.random() returned in call order, a_tape [Ta, B] = action_space.sample() values."""
        B = self.batch.num_envs
        u = torch.as_tensor(np.ascontiguousarray(u_tape, dtype=np.float64).reshape(-1, B)).to(self.device)
        a = torch.as_tensor(np.ascontiguousarray(a_tape, dtype=np.uint8).reshape(-1, B)).to(self.device)
        if a.shape[0] == 0:
            a = torch.zeros((1, B), dtype=torch.uint8, device=self.device)
        self._tapes = (u, a, torch.zeros((2, B), dtype=torch.int32, device=self.device))
        self._c = self._make_struct()

    # -- kernels
    def _stream(self):
        return cabi.current_stream(self.device)

    def act(self) -> torch.Tensor:
        rc = cabi.lib().maze_q_act(self.ctx.handle, C.byref(self.batch._c), C.byref(self._c), cabi.ptr(self.actions), self._stream())
        self.ctx.check(rc, "maze_q_act")
        return self.actions

    def learn(self):
        rc = cabi.lib().maze_q_update(self.ctx.handle, C.byref(self.batch._c), C.byref(self._c), self._stream())
        self.ctx.check(rc, "maze_q_update")

    def rollout(self, k_steps: int, mode: int = cabi.STEP_AUTORESET):
        """k_steps fused {get_action, env step, update} iterations per env in one launch."""
        rc = cabi.lib().maze_q_rollout(self.ctx.handle, C.byref(self.batch._c), C.byref(self._c), int(k_steps), int(mode),
                                       self._stream())
        self.ctx.check(rc, "maze_q_rollout")

    def check_overflow(self):
        if int(self.overflow.item()) != 0:
            raise cabi.MazeError(f"Q table of {self.capacity} rows overflowed: construct the agent with a larger capacity")

    # -- host views
    def table_host(self, which: str = "a") -> dict:
        """{(agent_id, agent_r, agent_c, target_r, target_c, best_dir_code): float64[4]} of the used rows."""
        keys = self.keys.cpu().numpy().view(np.uint64)
        vals = (self.q_a if which == "a" else self.q_b).cpu().numpy()
        used = np.nonzero(keys != np.uint64(cabi.Q_EMPTY))[0]
        out = {}
        for i in used:
            k = int(keys[i])
            out[(k >> 35, k & 0xff, (k >> 8) & 0xff, (k >> 16) & 0xff, (k >> 24) & 0xff, (k >> 32) & 7)] = vals[i].copy()
        return out


class _ReferenceNamedAgent:
    DOUBLE = False

    def __init__(self, env, learning_rate: float, initial_epsilon: float, epsilon_decay: float, final_epsilon: float,
                 discount_factor: float, eta: float, **kw):
        self.env = env
        batch = env.batch if hasattr(env, "batch") else env
        self.core = TabularAgent(batch, learning_rate, initial_epsilon, epsilon_decay, final_epsilon, discount_factor, eta,
                                 double_q=self.DOUBLE, **kw)
        self.lr = self.learning_rate = float(learning_rate)
        self.initial_epsilon, self.epsilon_decay, self.final_epsilon = initial_epsilon, epsilon_decay, final_epsilon
        self.eta = eta

    @property
    def discount_factor(self):
        return self.core.gamma

    @property
    def steps_done(self):
        return self.core.steps_done

    def get_action(self, obs=None):
        """Batched epsilon-greedy action for every env (obs is implicit: the env state on the device)."""
        return self.core.act()

    def update(self, obs=None, action=None, reward=None, terminated=None, next_obs=None):
        """Batched TD update for the transition env.step() just made with get_action()'s actions;
        also applies update_hyperparameter for the envs whose episode ended."""
        self.core.learn()

    def update_hyperparameter(self, is_better=None):
        """Applied per agent on the device at episode end (see update); kept for API parity."""

    def rollout(self, k_steps: int, mode=None):
        """k_steps fused {get_action, env.step, update} iterations per env in one launch.  `mode`
        defaults to the vector env's own step mode (autoreset, pool cycling); regeneration on win
        needs a generation launch between steps and is only available through get_action / update."""
        if mode is None:
            mode = getattr(self.env, "_mode", cabi.STEP_AUTORESET)
            if mode & cabi.STEP_WIN_QUEUE:
                raise cabi.MazeError("rollout() cannot regenerate mazes between its fused steps: use on_win='keep' / 'next', "
                                     "or the unfused get_action() / env.step() / update() loop")
        self.core.rollout(k_steps, mode)


class QAgent(_ReferenceNamedAgent):
    """agents/q_agent.py:8-79 on the device."""
    DOUBLE = False

    @property
    def q_values(self):
        return self.core.table_host("a")


class DQAgent(_ReferenceNamedAgent):
    """agents/dq_agent.py:5-73 on the device."""
    DOUBLE = True

    @property
    def q_a_values(self):
        return self.core.table_host("a")

    @property
    def q_b_values(self):
        return self.core.table_host("b")
