"""Host-side multi-GPU logic on CPU: index sharding and the statistics reduction over a
world_size-2 gloo group (the N>1 path of bench.py / MazeVectorEnv.episode_statistics)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maze_b200 import dist as mdist


@pytest.mark.parametrize("total,world", [(4096000, 8), (1000, 3), (7, 8), (0, 2), (65536, 1)])
def test_shard_range_partitions_the_index_space(total, world):
    shards = [mdist.shard_range(total, r, world) for r in range(world)]
    assert shards[0].start == 0 and shards[-1].stop == total
    for a, b in zip(shards[:-1], shards[1:]):
        assert a.stop == b.start
    assert max(s.count for s in shards) - min(s.count for s in shards) <= 1
    with pytest.raises(ValueError):
        mdist.shard_range(total, world, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_envs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, lr, w = mdist.env_from_torchrun()
        assert (r, lr, w) == (rank, rank, world)
        sh = mdist.shard_range(total_envs, r, w)
        # each rank "finishes" one episode per owned env, wins on even global ids, return = global id
        ids = torch.arange(sh.start, sh.stop, dtype=torch.float64)
        local = torch.tensor([ids.numel(), (ids % 2 == 0).sum().item(), (ids % 2 == 1).sum().item(), 10.0 * ids.numel(), ids.sum().item()],
                             dtype=torch.float64)
        red = mdist.reduce_statistics(local)
        out[rank] = mdist.statistics_dict(red)
    finally:
        dist.destroy_process_group()


def test_statistics_reduction_gloo_world2():
    world, total = 2, 1001
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
        res = dict(out)
    want = dict(episodes=total, wins=501, truncations=500, steps=10 * total, return_sum=float(total * (total - 1) // 2))
    for r in range(world):
        for k, v in want.items():
            assert res[r][k] == v, (r, k, res[r])
        assert res[r]["win_rate"] == pytest.approx(501 / 1001)


def test_reduce_is_identity_without_a_process_group():
    v = torch.arange(5, dtype=torch.float64)
    assert torch.equal(mdist.reduce_statistics(v), v)
    with pytest.raises(ValueError):
        mdist.reduce_statistics(torch.zeros(4, dtype=torch.float64))


@pytest.mark.parametrize("force_port", [False, True])
def test_reference_arm_of_the_bench_runs_without_a_gpu(force_port):
    """`bench.py --impl reference` is the CPU arm the driver runs beside ours: it must work on a box
    without a GPU, print exactly one JSON line with the contract keys, and honour its time budget.  With baseline/_ref
    present (tools/install_reference.py) it times the unmodified reference's own SimpleMazeEnv (kind "reference");
    without it -- or with MAZE_REF_FORCE_PORT -- the oracle's port (kind "port")."""
    import json
    import os
    import subprocess
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MAZE_REF_BUDGET_S="2")
    if force_port:
        env["MAZE_REF_FORCE_PORT"] = "1"
    have_ref = os.path.isfile(os.path.join(root, "baseline", "_ref", "gymnasium_env", "envs", "simple_maze_env.py"))
    want_kind = "reference" if have_ref and not force_port else "port"
    t0 = time.time()
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert set(d["config"]) == {"workload", "envs_per_gpu", "mazes_per_gpu", "l2", "parallelism"}      # the same keys as this repo's arm
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 3 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert time.time() - t0 < 120


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """The driver launches the reference arm like ours (torchrun, N ranks): rank 0 alone measures and prints,
    the other ranks exit 0 without work."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MAZE_REF_BUDGET_S="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(root, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
