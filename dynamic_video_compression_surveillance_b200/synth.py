"""Deterministic synthetic surveillance clips (arrays, not files).

Shared by the product benchmarks, the oracle and the tests, so that every
arm (CUDA path, CPU port, unmodified reference) sees byte-identical input.
Definition follows SURVEY.md section 8(d): a static random background (any
per-frame noise would make ``diff > 0`` everywhere at the reference's default
motion_threshold=0.5, frame_differencing.py:24,97) with four coloured
rectangles moving at fixed velocities and wrapping around.

Frame 0 only seeds ``prev_gray`` in the loop (frame_differencing.py:67-77).
"""
from __future__ import annotations

import numpy as np

_COLOURS = ((0, 0, 255), (0, 255, 0), (255, 0, 0), (255, 255, 255))   # BGR
_VELOCITY = ((3, 1), (-2, 2), (5, 0), (1, -3))

RESOLUTIONS = {
    "480p": (480, 640),
    "1080p": (1080, 1920),
    "4k": (2160, 3840),
}


class SyntheticClip:
    """Lazy frame source: ``clip[t]`` -> uint8 [H, W, 3] in BGR order."""

    def __init__(self, height: int, width: int, n_frames: int, seed: int = 0,
                 temporal_noise: bool = False, n_rects: int = 4):
        self.height, self.width, self.n_frames = int(height), int(width), int(n_frames)
        self.seed = int(seed)
        self.temporal_noise = bool(temporal_noise)
        rng = np.random.default_rng(self.seed)
        self.background = rng.integers(0, 256, (self.height, self.width, 3), dtype=np.uint8)
        self.rects = []
        for r in range(n_rects):
            w = max(2, min(self.width // 8 + 16 * r, self.width - 1))
            h = max(2, min(self.height // 8 + 8 * r, self.height - 1))
            x0 = int(rng.integers(0, self.width - w))
            y0 = int(rng.integers(0, self.height - h))
            self.rects.append((w, h, x0, y0, _VELOCITY[r % 4], _COLOURS[r % 4]))
        self._noise_seed = int(rng.integers(0, 2 ** 31 - 1))

    def __len__(self) -> int:
        return self.n_frames

    def rect_positions(self, t: int):
        out = []
        for (w, h, x0, y0, (vx, vy), _c) in self.rects:
            x = (x0 + vx * t) % (self.width - w)
            y = (y0 + vy * t) % (self.height - h)
            out.append((x, y, w, h))
        return out

    def render_into(self, t: int, out: np.ndarray) -> np.ndarray:
        np.copyto(out, self.background)
        for (x, y, w, h), rect in zip(self.rect_positions(t), self.rects):
            out[y:y + h, x:x + w] = rect[5]
        if self.temporal_noise:
            nrng = np.random.default_rng(self._noise_seed + t)
            noise = nrng.integers(-2, 3, out.shape, dtype=np.int16)
            np.copyto(out, np.clip(out.astype(np.int16) + noise, 0, 255).astype(np.uint8))
        return out

    def __getitem__(self, t: int) -> np.ndarray:
        if not 0 <= t < self.n_frames:
            raise IndexError(t)
        return self.render_into(t, np.empty((self.height, self.width, 3), np.uint8))

    def frames(self, start: int = 0, stop: int | None = None) -> np.ndarray:
        stop = self.n_frames if stop is None else stop
        out = np.empty((stop - start, self.height, self.width, 3), np.uint8)
        for i, t in enumerate(range(start, stop)):
            self.render_into(t, out[i])
        return out


def make_clip(resolution: str | tuple, n_frames: int, seed: int = 0, **kw) -> SyntheticClip:
    h, w = RESOLUTIONS[resolution] if isinstance(resolution, str) else resolution
    return SyntheticClip(h, w, n_frames, seed=seed, **kw)
