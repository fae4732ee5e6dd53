"""How far apart do runs of the SHARED tabular learner land?  (Concurrent writers of a Q-table row race, one wins: runs with the same seed
are not bit-identical.)  Prints late win rates and table differences of three runs; quoted by tests/test_gpu_qlearn.py."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ -> repo root
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import maze_b200 as mb
from maze_b200.agents import QAgent
from conftest import load_golden
z, _ = load_golden("qagent")
KW = dict(learning_rate=0.1, initial_epsilon=0.9, epsilon_decay=150, final_epsilon=0.05, discount_factor=0.7, eta=1e-2)
pool = mb.MazePool.from_grids([z["grid"]], [tuple(z["start"])], [tuple(z["goal"])], [False])
B = 2048
res = []
for run in range(3):
    batch = mb.MazeBatch(pool, B, stats=True)
    agent = QAgent(batch, envs_per_agent=B, seed=5, **KW)
    batch.reset()
    agent.rollout(1650)
    mid = batch.stats.cpu().numpy().copy()
    agent.rollout(300)
    late = batch.stats.cpu().numpy() - mid
    t = agent.core.table_host("a")
    res.append((late[1] / late[0], t))
    print("run", run, "late win rate", late[1] / late[0], "episodes", late[0], "keys", len(t))
for a in range(3):
    for b in range(a + 1, 3):
        ta, tb = res[a][1], res[b][1]
        keys = set(ta) & set(tb)
        d = np.array([np.abs(ta[k] - tb[k]).max() for k in keys])
        mag = np.array([np.abs(ta[k]).max() for k in keys])
        same_argmax = np.mean([int(np.argmax(ta[k]) == np.argmax(tb[k])) for k in keys])
        print(a, b, "common keys", len(keys), "of", len(ta), len(tb), "identical", (d == 0).mean(), "median |dq|", np.median(d), "p99", np.quantile(d, 0.99), "max", d.max(), "median |q|", np.median(mag), "greedy action agrees", same_argmax)
