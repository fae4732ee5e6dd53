"""BASELINE.json configs[0] with the reference's OWN consumer code: the unmodified agents/q_agent.py (dq_agent.py) and
lib/trainers/off_policy_trainer.py, loaded from baseline/_ref (tools/install_reference.py copies the reference tree
there; git-ignored, shipped to the GPU box with the snapshot), drive this repo's
gymnasium_env.envs.simple_maze_env.SimpleMazeEnv exactly as training_examples/euclidean_mazes/costant_sizes/test_q.py:27-49
does -- RecordEpisodeStatistics wrapper, isinstance checks, update_maze() on every win, ComplexityEvaluation of the
won maze included.  Every transition the trainer saw is then replayed through the oracle (bit-exact step check), and
the reference agent's Q table is compared with oracle.qlearn fed the recorded draws."""
import logging
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _load(relpath, name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref_modules():
    if not os.path.exists(os.path.join(REF, "agents", "q_agent.py")):
        pytest.skip("baseline/_ref is missing: run tools/install_reference.py in the build container")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_shim
    try:
        import gymnasium  # noqa: F401
    except Exception:
        ref_shim._install_gymnasium()      # stand-in for `import gymnasium as gym` in the reference files
    # `gymnasium_env.envs.*` and `lib.maze_difficulty_evaluation.*` resolve to THIS repo's classes (conftest puts the
    # package directory first on sys.path): that is the drop-in; only the consumer files come from the reference.
    import gymnasium_env.envs.simple_maze_env as ours
    assert "maze-solving-agent-gymnasium_b200" in ours.__file__
    trainer = _load("lib/trainers/off_policy_trainer.py", "_ref_off_policy_trainer")
    q = _load("agents/q_agent.py", "_ref_q_agent")
    dq = _load("agents/dq_agent.py", "_ref_dq_agent")
    return dict(OffPolicyTrainer=trainer.OffPolicyTrainer, QAgent=q.QAgent, DQAgent=dq.DQAgent, SimpleMazeEnv=ours.SimpleMazeEnv)


class _Recorder:
    """Records what the trainer sees; forwards everything else to the env (the trainer reaches env.env.*)."""

    def __init__(self, env):
        self.env = env
        self.episodes = []

    def reset(self, **kw):
        obs, info = self.env.reset(**kw)
        e = self.env
        self.episodes.append(dict(grid=np.array(e.maze_map, dtype=np.uint8), start=tuple(int(x) for x in e._start_pos),
                                  goal=tuple(int(x) for x in e._target_location), max_steps=int(e.max_steps_taken), obs0=obs, steps=[]))
        return obs, info

    def step(self, action):
        out = self.env.step(action)
        self.episodes[-1]["steps"].append((int(action), out))
        return out

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.env, name)


def _same_obs(a, b):
    return all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) and np.asarray(a[k]).dtype == np.asarray(b[k]).dtype for k in ("agent", "target", "best dir"))


@pytest.mark.parametrize("agent_name", ["QAgent", "DQAgent"])
def test_reference_trainer_and_agent_run_unmodified_on_the_device_env(ref_modules, agent_name):
    import gymnasium as gym
    from oracle.env_port import ClosedFormEnv
    from oracle.qlearn import OracleQAgent
    import random
    random.seed(123)
    np.random.seed(123)
    shape, n_episodes = (21, 21), 30
    env = ref_modules["SimpleMazeEnv"](shape)
    rec = _Recorder(env)
    wrapped = gym.wrappers.RecordEpisodeStatistics(rec, buffer_length=n_episodes)     # test_q.py:28
    assert wrapped.env is rec
    kw = dict(env=wrapped, learning_rate=1e-1, initial_epsilon=0.95, final_epsilon=0.05, epsilon_decay=shape[0] * shape[1] // 2, eta=1e-2)
    kw["discount_factor"] = 0.7       # test_dq.py:37 passes `gamma=`, a TypeError in the reference
    agent = ref_modules[agent_name](**kw)
    # record the agent's draws so that the oracle agent can consume the same numbers
    draws_u, draws_a = [], []
    real_random, real_sample = np.random.random, wrapped.action_space.sample

    def rec_random(*a, **k):
        v = real_random(*a, **k)
        draws_u.append(float(v))
        return v

    def rec_sample():
        v = real_sample()
        draws_a.append(int(v))
        return v

    np.random.random = rec_random
    env.action_space.sample = rec_sample
    try:
        trainer = ref_modules["OffPolicyTrainer"](wrapped, agent, logging.getLogger("ref_consumer"))
        trainer.train(n_episodes)                                                      # test_q.py:49
    finally:
        np.random.random = real_random
        env.action_space.sample = real_sample

    # ---- (a) every transition the reference trainer consumed equals the oracle's on the same maze and actions
    episodes = [e for e in rec.episodes if e["steps"]]
    assert len(episodes) == n_episodes
    n_steps = wins = 0
    for ep in episodes:
        o = ClosedFormEnv(ep["grid"], ep["start"], ep["goal"], False)
        assert o.max_steps == ep["max_steps"]
        obs, _ = o.reset()
        assert _same_obs(obs, ep["obs0"])
        for action, (nobs, reward, truncated, terminated, info) in ep["steps"]:
            oobs, r, otr, ote, oinfo = o.step(action)
            assert _same_obs(oobs, nobs), (action, oobs, nobs)
            assert np.float64(r).view(np.uint64) == np.float64(reward).view(np.uint64)
            assert bool(otr) == bool(truncated) and bool(ote) == bool(terminated)
            assert oinfo["distance"] == info["distance"]
            n_steps += 1
        wins += int(bool(ep["steps"][-1][1][3]))
    assert n_steps > 500
    assert len(env.mazes) == 1 + wins            # update_maze() after every win appended a maze (test_q.py:52 takes len)

    # ---- (b) the reference agent's table equals the oracle agent's after the same transitions and draws
    from oracle.qlearn import Draws, obs_key, parse_reference_key
    double = agent_name == "DQAgent"
    ora = OracleQAgent(learning_rate=1e-1, initial_epsilon=0.95, epsilon_decay=shape[0] * shape[1] // 2, final_epsilon=0.05,
                       discount_factor=0.7, eta=1e-2, draws=Draws(draws_u, draws_a), double_q=double)
    for ep in episodes:
        obs, cum = ep["obs0"], 0.0
        for action, (nobs, reward, truncated, terminated, _) in ep["steps"]:
            a = ora.get_action(obs_key(obs))
            assert a == action
            ora.update(obs_key(obs), a, reward, terminated, obs_key(nobs))
            cum += reward
            obs = nobs
        ora.update_hyperparameter(cum > 0)       # off_policy_trainer.py:77-79 (prev_cum_rew is reset to 0 every episode)
    tables = [(agent.q_values, ora.q_a)] if not double else [(agent.q_a_values, ora.q_a), (agent.q_b_values, ora.q_b)]
    for ref_tab, ora_tab in tables:
        ref_tab = {parse_reference_key(k): v for k, v in ref_tab.items()}
        assert set(ref_tab.keys()) == set(ora_tab.keys())
        for k, v in ref_tab.items():
            np.testing.assert_array_equal(np.asarray(v, dtype=np.float64).view(np.uint64), np.asarray(ora_tab[k], dtype=np.float64).view(np.uint64))
    assert agent.discount_factor == ora.discount_factor
    assert ora.draws.iu == len(draws_u) and ora.draws.ia == len(draws_a)
