"""Tabular Q-learning / double Q-learning oracle (test infrastructure; see oracle/__init__.py).

Restates agents/q_agent.py:8-79 (QAgent) and agents/dq_agent.py:5-73 (DQAgent) plus the episode
loop of lib/trainers/off_policy_trainer.py:38-51,76-78, with the random draws supplied by the
caller (the reference draws from numpy's global RNG; the goldens recorded those draws).  Tables
are dicts keyed by the integer content of the observation instead of str(obs).  Pinned against
tests/golden/qagent.npz (tables of the unmodified reference agents after 12 episodes).
"""
from __future__ import annotations

import math
import re

import numpy as np


def obs_key(obs):
    """(agent_r, agent_c, target_r, target_c, best_dr, best_dc) -- what str(obs) encodes (q_agent.py:54)."""
    return (int(obs["agent"][0]), int(obs["agent"][1]), int(obs["target"][0]), int(obs["target"][1]),
            int(obs["best dir"][0]), int(obs["best dir"][1]))


def parse_reference_key(s: str):
    """Invert str(obs) of the reference's observation dict."""
    nums = [int(x) for x in re.findall(r"-?\d+", s.replace("int32", "").replace("int64", ""))]
    assert len(nums) == 6, s
    return tuple(nums)


class Draws:
    """Sequential source of np.random.random() / action_space.sample() values."""

    def __init__(self, u, a):
        self.u, self.a = list(u), list(a)
        self.iu = self.ia = 0

    def random(self):
        v = self.u[self.iu]
        self.iu += 1
        return v

    def sample(self):
        v = self.a[self.ia]
        self.ia += 1
        return int(v)


class OracleQAgent:
    def __init__(self, learning_rate, initial_epsilon, epsilon_decay, final_epsilon, discount_factor, eta, draws, double_q=False):
        self.lr, self.eta = learning_rate, eta
        self.initial_epsilon, self.epsilon_decay, self.final_epsilon = initial_epsilon, epsilon_decay, final_epsilon
        self.discount_factor = discount_factor
        self.draws = draws
        self.double_q = double_q
        self.q_a, self.q_b = {}, {}
        self.steps_done = 0

    def _row(self, table, key):
        if key not in table:
            table[key] = np.zeros(4)
        return table[key]

    def epsilon(self):
        # q_agent.py:49
        return self.final_epsilon + (self.initial_epsilon - self.final_epsilon) * math.exp(-1. * self.steps_done / self.epsilon_decay)

    def get_action(self, key):
        eps = self.epsilon()
        self.steps_done += 1
        if self.draws.random() < eps:
            return self.draws.sample()
        return int(np.argmax(self._row(self.q_a, key)))

    def update(self, key, action, reward, terminated, next_key):
        if not self.double_q:   # q_agent.py:56-72
            future = (not terminated) * np.max(self._row(self.q_a, next_key))
            row = self._row(self.q_a, key)
            td = reward + self.discount_factor * future - row[action]
            row[action] = row[action] + self.lr * td
            return
        # dq_agent.py:49-66
        if self.draws.random() < 0.5:
            best = self.get_action(next_key)
            row = self._row(self.q_a, key)
            td = reward + self.discount_factor * self._row(self.q_b, next_key)[best] - row[action]
            row[action] = row[action] + self.lr * td
        else:
            best = self.get_action(next_key)
            row = self._row(self.q_b, key)
            td = reward + self.discount_factor * self._row(self.q_a, next_key)[best] - row[action]
            row[action] = row[action] + self.lr * td

    def update_hyperparameter(self, is_better):
        self.discount_factor = self.discount_factor + self.eta if is_better else self.discount_factor - self.eta


def run_episodes(env, agent, n_episodes=None, n_steps=None):
    """The loop of off_policy_trainer.py:29-51,76-78 over an oracle env (reference return order).
    Stops after n_episodes episodes or n_steps transitions.  Returns the transition log."""
    log = dict(action=[], reward=[], term=[], trunc=[], gamma=[])
    ep = 0
    while n_episodes is None or ep < n_episodes:
        obs, _ = env.reset()
        done, cum = False, 0
        while not done:
            a = agent.get_action(obs_key(obs))
            nobs, r, trunc, term, _ = env.step(a)
            log["action"].append(a); log["reward"].append(float(r)); log["term"].append(bool(term)); log["trunc"].append(bool(trunc))
            log["gamma"].append(float(agent.discount_factor))
            agent.update(obs_key(obs), a, r, term, obs_key(nobs))
            cum += r
            done = term or trunc
            obs = nobs
            if n_steps is not None and len(log["action"]) >= n_steps:
                return log
        agent.update_hyperparameter(cum > 0)
        ep += 1
    return log
