"""Vector-of-envs driver for the oracle (test infrastructure): gymnasium next-step autoreset
semantics layered over single-env oracles, mirroring what maze_step does with
MAZE_STEP_AUTORESET (and MAZE_STEP_WIN_NEXT for pool cycling)."""
from __future__ import annotations

import numpy as np

from .env_port import ClosedFormEnv, MazeTables


class OracleVector:
    def __init__(self, mazes, env_maze, autoreset=True, win_next=False, pool_stride=1, env_cls=ClosedFormEnv):
        """mazes: list of dict(grid, start, goal, toroidal); env_maze: maze index per env."""
        self.mazes = mazes
        self.tables = [None] * len(mazes)
        self.env_maze = list(int(m) for m in env_maze)
        self.autoreset, self.win_next, self.pool_stride = autoreset, win_next, pool_stride
        self.env_cls = env_cls
        self.envs = [self._make(m) for m in self.env_maze]
        self.pending = [False] * len(self.envs)
        self.won = [False] * len(self.envs)

    def _make(self, m):
        mz = self.mazes[m]
        if self.env_cls is ClosedFormEnv:
            if self.tables[m] is None:
                self.tables[m] = MazeTables(mz["grid"], mz["start"], mz["goal"], mz["toroidal"])
            return ClosedFormEnv(mz["grid"], mz["start"], mz["goal"], mz["toroidal"], tables=self.tables[m])
        return self.env_cls(mz["grid"], mz["start"], mz["goal"], mz["toroidal"])

    def reset(self):
        out = [e.reset() for e in self.envs]
        self.pending = [False] * len(self.envs)
        return self._pack([o[0] for o in out], np.zeros(len(out)), np.zeros(len(out), bool), np.zeros(len(out), bool))

    def step(self, actions):
        obs, rew, term, trunc = [], [], [], []
        for i, (env, a) in enumerate(zip(self.envs, actions)):
            if self.autoreset and self.pending[i]:
                if self.win_next and self.won[i]:
                    self.env_maze[i] = (self.env_maze[i] + self.pool_stride) % len(self.mazes)
                    env = self.envs[i] = self._make(self.env_maze[i])
                o, _ = env.reset()
                r, tr, te = 0.0, False, False
                self.pending[i] = False
            else:
                o, r, tr, te, _ = env.step(int(a))
                self.pending[i] = bool(tr or te)
                self.won[i] = bool(te)
            obs.append(o); rew.append(float(r)); term.append(te); trunc.append(tr)
        return self._pack(obs, np.array(rew), np.array(term), np.array(trunc))

    @staticmethod
    def _pack(obs, rew, term, trunc):
        return dict(agent=np.stack([o["agent"] for o in obs]).astype(np.int64),
                    target=np.stack([o["target"] for o in obs]).astype(np.int64),
                    best=np.stack([o["best dir"] for o in obs]).astype(np.int64),
                    reward=np.asarray(rew, dtype=np.float64), term=np.asarray(term, bool), trunc=np.asarray(trunc, bool))
