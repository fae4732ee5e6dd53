"""oracle/render.py against frames of the unmodified lib/maze_view.py (tests/golden/render.npz, made by
make_golden_render.py on a software pygame: the reference draws rectangles only, which rasterise exactly)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.render import Canvas


def test_canvas_reproduces_the_reference_frames():
    z = np.load(f"{GOLDEN}/render.npz")
    total = 0
    for k in range(int(z["count"])):
        grid, start = z[f"grid{k}"], z[f"start{k}"]
        frames, moved, pos = z[f"frames{k}"], z[f"moved{k}"], z[f"pos{k}"]
        c = Canvas(grid, start)
        assert frames[0].shape == (grid.shape[0] * 16, grid.shape[1] * 16, 3)
        np.testing.assert_array_equal(c.frame(), frames[0], err_msg=f"maze {k}: frame after construction")
        for t in range(len(moved)):
            if moved[t]:
                c.move_to(pos[t])
            np.testing.assert_array_equal(c.frame(), frames[1 + t], err_msg=f"maze {k} step {t}")
            total += 1
        assert moved.any() and not moved.all()      # both a successful and a blocked move are in the walk
        c.move_to(start)                            # _reset_agent (maze_view.py:154-158): trail on the block left, agent at the start
        np.testing.assert_array_equal(c.frame(), frames[-1], err_msg=f"maze {k}: frame after _reset_agent")
    assert total == 240


def test_fixture_is_what_the_reference_view_draws_today():
    """Where /root/reference exists (the build container), run its lib/maze_view.py again on the software pygame and compare
    with the committed frames: the fixture is the reference's output, not a copy of the oracle's."""
    sys.path.insert(0, GOLDEN)
    import pygame_raster
    from ref_shim import REFERENCE_ROOT, reference_available
    if not reference_available():
        pytest.skip("/root/reference is not present on this machine")
    saved = sys.modules.get("pygame")
    pygame_raster.install()
    try:
        spec = importlib.util.spec_from_file_location("ref_maze_view_check", os.path.join(REFERENCE_ROOT, "lib", "maze_view.py"))
        mv = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mv)
        z = np.load(f"{GOLDEN}/render.npz")
        actions = ((1, 0), (-1, 0), (0, 1), (0, -1))
        for k in (0, 4):   # one bordered, one toroidal maze
            cls = mv.ToroidalMazeView if bool(z[f"toroidal{k}"]) else mv.SimpleMazeView
            grid = z[f"grid{k}"]
            view = cls(maze_map=grid.tolist(), start_position=tuple(int(v) for v in z[f"start{k}"]),
                       goal_position=tuple(int(v) for v in z[f"goal{k}"]), maze_size=grid.shape)
            np.testing.assert_array_equal(view.view_update("rgb_array"), z[f"frames{k}"][0])
            for t, a in enumerate(z[f"actions{k}"]):
                assert bool(view.move_agent(actions[int(a)])) == bool(z[f"moved{k}"][t])
                np.testing.assert_array_equal(view.view_update("rgb_array"), z[f"frames{k}"][1 + t], err_msg=f"maze {k} step {t}")
    finally:
        if saved is not None:
            sys.modules["pygame"] = saved
        else:
            sys.modules.pop("pygame", None)
