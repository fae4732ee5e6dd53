"""Q / double-Q oracle vs the tables the unmodified reference agents ended with (qagent.npz)."""
import numpy as np
import pytest

from oracle.env_port import ClosedFormEnv
from oracle.qlearn import Draws, OracleQAgent, parse_reference_key, run_episodes

KW = dict(learning_rate=0.1, initial_epsilon=0.9, epsilon_decay=150, final_epsilon=0.05, discount_factor=0.7, eta=1e-2)


@pytest.mark.parametrize("name", ["q", "dq"])
def test_replay_reproduces_reference_tables(golden_qagent, name):
    z, _ = golden_qagent
    env = ClosedFormEnv(z["grid"], tuple(z["start"]), tuple(z["goal"]), False)
    assert env.max_steps == int(z["max_steps"])
    draws = Draws(z[f"{name}_u"], z[f"{name}_a"])
    agent = OracleQAgent(draws=draws, double_q=(name == "dq"), **KW)
    log = run_episodes(env, agent, n_episodes=12)
    np.testing.assert_array_equal(log["action"], z[f"{name}_action"])
    np.testing.assert_array_equal(np.array(log["reward"]).view(np.uint64), z[f"{name}_reward"].view(np.uint64))
    np.testing.assert_array_equal(log["term"], z[f"{name}_term"])
    np.testing.assert_array_equal(np.array(log["gamma"]).view(np.uint64), z[f"{name}_gamma"].view(np.uint64))
    assert draws.iu == len(z[f"{name}_u"]) and draws.ia == len(z[f"{name}_a"])
    assert agent.steps_done == int(z[f"{name}_steps_done"])
    assert agent.discount_factor == float(z[f"{name}_gamma_final"])
    for ti, table in enumerate([agent.q_a] if name == "q" else [agent.q_a, agent.q_b]):
        keys = [parse_reference_key(str(k)) for k in z[f"{name}_tab{ti}_keys"]]
        vals = z[f"{name}_tab{ti}_vals"]
        assert set(keys) == set(table)
        for k, v in zip(keys, vals):
            np.testing.assert_array_equal(table[k].view(np.uint64), v.view(np.uint64), err_msg=str(k))
