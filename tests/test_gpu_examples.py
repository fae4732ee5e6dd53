"""The shipped example (configs[4]: DDQN loop over the device env, replay and action selection)
runs end to end and reports sane numbers."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_ddqn_example_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "train_ddqn.py"), "--envs", "512", "--iters", "40",
                          "--batch", "256", "--shape", "21", "--memory", "65536"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["env_steps_per_s"] > 0 and line["optimizer_steps_per_s"] > 0 and line["final_loss"] is not None
    assert line["episodes"] > 0          # 21x21 mazes: episodes end within 40 steps
