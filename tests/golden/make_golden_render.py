"""render.npz: frames of the UNMODIFIED lib/maze_view.py (build container only; test infrastructure).

The reference's views run on the software pygame of pygame_raster.py (rectangles only: exact).  For a few mazes of both
topologies: the frame after construction, and the frames after a scripted walk (successful moves, blocked moves, a wrap on
the torus, a reset) -- `view_update("rgb_array")`, i.e. what BaseMazeEnv.render returns (base_maze_env.py:212-215).
Usage: python tests/golden/make_golden_render.py   (writes tests/golden/render.npz)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pygame_raster  # noqa: E402

pygame_raster.install()
from ref_shim import REFERENCE_ROOT  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_maze_view", os.path.join(REFERENCE_ROOT, "lib", "maze_view.py"))
mv = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mv)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import random  # noqa: E402

from oracle.generation import gen_maze, gen_maze_no_border  # noqa: E402

ACTIONS = ((1, 0), (-1, 0), (0, 1), (0, -1))   # base_maze_env.py: down, up, right, left


def walk(view_cls, grid, start, goal, toroidal, seed):
    """Frames: [0] after construction, then one per scripted action (moved or not), then one after _reset_agent."""
    rng = np.random.default_rng(seed)
    # positions as TUPLES, like the envs pass them (gen_maze's return values, simple_maze_env.py:26-32): with a numpy start
    # position SimpleMazeView.move_agent's in-place `+=` would walk start_position along with the agent
    view = view_cls(maze_map=grid.tolist(), start_position=tuple(int(v) for v in start), goal_position=tuple(int(v) for v in goal),
                    maze_size=grid.shape)
    frames = [view.view_update("rgb_array")]
    acts, moved, pos = [], [], []
    for _ in range(40):
        a = int(rng.integers(0, 4))
        ok = bool(view.move_agent(ACTIONS[a]))
        frames.append(view.view_update("rgb_array"))
        acts.append(a)
        moved.append(ok)
        pos.append(tuple(int(v) for v in view._agent_position))
    view._reset_agent()
    frames.append(view.view_update("rgb_array"))
    return np.stack(frames), np.array(acts, np.uint8), np.array(moved, np.uint8), np.array(pos, np.int32)


def main():
    out = {}
    k = 0
    for toroidal, cls in ((False, mv.SimpleMazeView), (True, mv.ToroidalMazeView)):
        for shape, algo, seed in (((9, 9), "r-prim", 1), ((11, 15), "dfs", 2), ((15, 11), "prim&kill", 3)):
            start, goal, grid = (gen_maze_no_border if toroidal else gen_maze)(shape, algo, random.Random(seed))
            frames, acts, moved, pos = walk(cls, np.asarray(grid), start, goal, toroidal, 100 + seed)
            out[f"grid{k}"], out[f"start{k}"], out[f"goal{k}"] = np.asarray(grid, np.uint8), np.array(start, np.int32), np.array(goal, np.int32)
            out[f"toroidal{k}"] = np.array(toroidal)
            out[f"frames{k}"], out[f"actions{k}"], out[f"moved{k}"], out[f"pos{k}"] = frames, acts, moved, pos
            k += 1
    out["count"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out)
    print("wrote render.npz:", k, "mazes,", sum(out[f"frames{i}"].shape[0] for i in range(k)), "frames")


if __name__ == "__main__":
    main()
