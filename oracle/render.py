"""Frame oracle (test infrastructure; see oracle/__init__.py): the drawing calls of
lib/maze_view.py replayed on a numpy canvas.

Parity PINNED (round 2): tests/test_oracle_render.py compares this canvas with 252 frames of the UNMODIFIED
lib/maze_view.py (tests/golden/render.npz).  pygame is not installed in the build container, so the reference's
views were driven on tests/golden/pygame_raster.py, a software pygame for the calls the renderer makes -- it draws
axis-aligned rectangles only (fill and one-pixel outline), which rasterise unambiguously.  The rules: __draw_maze
:88-96 (tile fill + (59, 66, 82) outline), _draw_agent :98-104 (8 x 8 square at offset 4), _draw_cell :148-152
(fill + (208, 135, 112) outline on the block the agent leaves), move_agent :167-180 / :184-197, _reset_agent
:154-158, and the layers being one pixel smaller than the window (:41-44).
"""
from __future__ import annotations

import numpy as np

TILE = 16
CELL_COLORS = ((46, 52, 64), (236, 239, 244), (163, 190, 140))
AGENT_COLOR = (94, 129, 172)
OUTLINE, TRAIL = (59, 66, 82), (208, 135, 112)


class Canvas:
    """Event-driven: draw the maze once, then replay agent moves as the view would."""

    def __init__(self, grid, start):
        self.g = np.asarray(grid)
        H, W = self.g.shape
        self.layer = np.zeros((H * TILE, W * TILE, 3), dtype=np.uint8)
        for r in range(H):
            for c in range(W):
                self._cell(r, c, OUTLINE)
        self.pos = (int(start[0]), int(start[1]))
        self._agent()

    def _cell(self, r, c, outline):
        y, x = r * TILE, c * TILE
        self.layer[y:y + TILE, x:x + TILE] = outline
        self.layer[y + 1:y + TILE - 1, x + 1:x + TILE - 1] = CELL_COLORS[int(self.g[r, c])]

    def _agent(self):
        y, x = self.pos[0] * TILE + TILE // 4, self.pos[1] * TILE + TILE // 4
        self.layer[y:y + TILE // 2, x:x + TILE // 2] = AGENT_COLOR

    def move_to(self, new_pos):
        """A successful move_agent: repaint the block being left with the trail outline, draw the agent."""
        self._cell(self.pos[0], self.pos[1], TRAIL)
        self.pos = (int(new_pos[0]), int(new_pos[1]))
        self._agent()

    def frame(self):
        out = self.layer.copy()
        out[-1, :] = 0      # the layers are (w - 1, h - 1): the window's last row / column keep the background
        out[:, -1] = 0
        return out
