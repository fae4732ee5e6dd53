"""A software `pygame` for the few calls lib/maze_view.py makes (build container only; test infrastructure).

pygame is not installed here, but the reference's renderer only uses axis-aligned rectangles on RGBA surfaces, which
rasterise unambiguously: `draw.rect(surface, color, Rect(x, y, w, h), 0)` fills [x, x + w) x [y, y + h) with the colour
(alpha included, no blending, clipped to the surface); width 1 draws that rectangle's one-pixel border; `blit` of a
per-pixel-alpha surface blends with its alpha (only 0 and 255 occur here); `surfarray.array3d` returns [x, y, rgb].  With
this module installed as `pygame`, the UNMODIFIED lib/maze_view.py runs and yields the frames a real pygame would give,
which pin oracle/render.py and maze_render (make_golden_render.py -> render.npz).
"""
from __future__ import annotations

import types

import numpy as np

QUIT = 256


class Rect:
    def __init__(self, x, y, w, h):
        self.x, self.y, self.w, self.h = int(x), int(y), int(w), int(h)


class Surface:
    def __init__(self, size, alpha=False):
        w, h = int(size[0]), int(size[1])
        self.px = np.zeros((h, w, 4), dtype=np.uint8)   # [y, x, rgba]; a fresh surface is black
        self.px[:, :, 3] = 0 if alpha else 255
        self.per_pixel_alpha = alpha

    def convert(self):
        s = Surface((self.px.shape[1], self.px.shape[0]), alpha=False)
        s.px[:, :, :3] = self.px[:, :, :3]
        return s

    def convert_alpha(self):
        # pygame.Surface(size).convert_alpha(): the black, opaque pixels of the new surface keep alpha 255
        s = Surface((self.px.shape[1], self.px.shape[0]), alpha=True)
        s.px[:] = self.px
        return s

    def blit(self, src, dest):
        dx, dy = int(dest[0]), int(dest[1])
        h = min(src.px.shape[0], self.px.shape[0] - dy)
        w = min(src.px.shape[1], self.px.shape[1] - dx)
        s = src.px[:h, :w].astype(np.uint32)
        d = self.px[dy:dy + h, dx:dx + w].astype(np.uint32)
        a = s[:, :, 3:4] if src.per_pixel_alpha else np.full_like(s[:, :, 3:4], 255)
        # SDL's blend: dst + ((src - dst) * a >> 8), exact at a = 0; opaque pixels copy
        out = np.where(a == 255, s[:, :, :3], (d[:, :, :3] * (255 - a) + s[:, :, :3] * a) // 255)
        self.px[dy:dy + h, dx:dx + w, :3] = out.astype(np.uint8)


def _rect(surface, color, rect, width=0):
    c = tuple(int(v) for v in color) + ((255,) if len(color) == 3 else ())
    H, W = surface.px.shape[:2]
    x0, y0, x1, y1 = max(rect.x, 0), max(rect.y, 0), min(rect.x + rect.w, W), min(rect.y + rect.h, H)
    if x0 >= x1 or y0 >= y1:
        return
    if width == 0:
        surface.px[y0:y1, x0:x1] = c
        return
    assert width == 1, "the reference only draws one-pixel outlines"
    for (ya, yb, xa, xb) in ((rect.y, rect.y + 1, rect.x, rect.x + rect.w), (rect.y + rect.h - 1, rect.y + rect.h, rect.x, rect.x + rect.w),
                             (rect.y, rect.y + rect.h, rect.x, rect.x + 1), (rect.y, rect.y + rect.h, rect.x + rect.w - 1, rect.x + rect.w)):
        ya, yb, xa, xb = max(ya, 0), min(yb, H), max(xa, 0), min(xb, W)
        if ya < yb and xa < xb:
            surface.px[ya:yb, xa:xb] = c


class _Display:
    def __init__(self):
        self.surface = None

    def set_caption(self, *_):
        pass

    def init(self):
        pass

    def set_mode(self, size):
        self.surface = Surface(size)
        return self.surface

    def get_surface(self):
        return self.surface

    def flip(self):
        pass

    def update(self):
        pass

    def quit(self):
        pass


def install():
    """Put the stub into sys.modules as `pygame` (before the reference is imported)."""
    import sys
    pg = types.ModuleType("pygame")
    pg.QUIT = QUIT
    pg.Rect, pg.Surface = Rect, Surface
    pg.init = lambda: None
    pg.quit = lambda: None
    pg.display = _Display()
    pg.draw = types.SimpleNamespace(rect=_rect)
    pg.event = types.SimpleNamespace(get=lambda: [])
    pg.surfarray = types.SimpleNamespace(array3d=lambda s: np.ascontiguousarray(s.px[:, :, :3].transpose(1, 0, 2)))
    pg.time = types.SimpleNamespace(Clock=lambda: types.SimpleNamespace(tick=lambda *_: None))
    sys.modules["pygame"] = pg
    return pg
