"""Late win rates of the shared tabular learner over several same-seed runs, and of the never-learning baseline (lr = 0):
the numbers behind the bounds of tests/test_gpu_qlearn.py::test_shared_learner_runs_differ_within_bounds."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "maze-solving-agent-gymnasium_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import maze_b200 as mb  # noqa: E402
from conftest import load_golden  # noqa: E402
from maze_b200.agents import QAgent  # noqa: E402

KW = dict(learning_rate=0.1, initial_epsilon=0.9, epsilon_decay=150, final_epsilon=0.05, discount_factor=0.7, eta=1e-2)
z, _ = load_golden("qagent")
pool = mb.MazePool.from_grids([z["grid"]], [tuple(z["start"])], [tuple(z["goal"])], [False])
B = 2048


def late_rate(lr):
    batch = mb.MazeBatch(pool, B, stats=True)
    agent = QAgent(batch, envs_per_agent=B, seed=5, **dict(KW, learning_rate=lr))
    batch.reset()
    agent.rollout(1650)
    mid = batch.stats.cpu().numpy().copy()
    agent.rollout(300)
    late = batch.stats.cpu().numpy() - mid
    return float(late[1] / late[0])


print("baseline (lr = 0):", [round(late_rate(0.0), 3) for _ in range(3)])
print("shared learner (lr = 0.1):", [round(late_rate(0.1), 3) for _ in range(12)])
