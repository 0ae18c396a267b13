"""Deterministic synthetic inputs (simulator texture + scripted camera path).

The reference ships no texture image and no test clip, and its
simulator is keyboard driven (/root/reference/src/main_utils.cpp:327-369) with
steps far too coarse for a benchmark.  This module provides what the survey
(§8d) specifies instead:

* a seeded, corner-rich procedural floor texture (multi-octave value noise at
  half contrast + random filled rotated rectangles), and
* a seeded scripted camera path (drift + sinusoid + Gaussian jitter) around the
  default pose of /root/reference/src/main.cpp:29-36.

Only numpy is used so the exact same bytes are produced on every machine
(numpy's PCG64 streams are platform independent).  Nothing here is on the
hot path (input generation only).
"""
from __future__ import annotations

import os
import tempfile

import numpy as np

TEXTURE_SEED = 7
PATH_SEED = 0


def _value_noise(rng: np.random.Generator, size: int, cells: int) -> np.ndarray:
    """Bilinear-interpolated lattice noise in [0,1), periodic (tileable)."""
    lattice = rng.random((cells, cells))
    t = np.arange(size, dtype=np.float64) * (cells / size)
    i0 = np.floor(t).astype(np.int64) % cells
    i1 = (i0 + 1) % cells
    f = t - np.floor(t)
    f = f * f * (3.0 - 2.0 * f)
    top = lattice[i0][:, i0] * (1 - f)[None, :] + lattice[i0][:, i1] * f[None, :]
    bot = lattice[i1][:, i0] * (1 - f)[None, :] + lattice[i1][:, i1] * f[None, :]
    return top * (1 - f)[:, None] + bot * f[:, None]


def _fill_rot_rect(img: np.ndarray, cx: float, cy: float, w: float, h: float,
                   ang: float, color: np.ndarray) -> None:
    """Fill a rotated rectangle by testing pixel centres (pure numpy, exact)."""
    size = img.shape[0]
    r = 0.5 * np.hypot(w, h) + 1.0
    x0, x1 = int(max(0, np.floor(cx - r))), int(min(size, np.ceil(cx + r) + 1))
    y0, y1 = int(max(0, np.floor(cy - r))), int(min(size, np.ceil(cy + r) + 1))
    if x0 >= x1 or y0 >= y1:
        return
    ys, xs = np.mgrid[y0:y1, x0:x1]
    dx = xs - cx
    dy = ys - cy
    c, s = np.cos(ang), np.sin(ang)
    u = dx * c + dy * s
    v = -dx * s + dy * c
    m = (np.abs(u) <= 0.5 * w) & (np.abs(v) <= 0.5 * h)
    img[y0:y1, x0:x1][m] = color


def make_texture(size: int = 2048, seed: int = TEXTURE_SEED,
                 n_rects: int | None = None, cache: bool = True) -> np.ndarray:
    """Seeded corner-rich BGR u8 texture, `size` x `size` x 3.

    SURVEY.md §8d: "multi-octave value noise at half contrast + ~17 k random
    filled rotated rectangles at scales 10-160 px" for 2048^2.  `n_rects`
    scales with the area when the texture is smaller (unit tests).
    """
    if n_rects is None:
        n_rects = int(17000 * (size / 2048.0) ** 2)
    cache_path = os.path.join(tempfile.gettempdir(),
                              f"vstab_texture_{size}_{seed}_{n_rects}.npy")
    if cache and os.path.exists(cache_path):
        try:
            tex = np.load(cache_path)
            if tex.shape == (size, size, 3) and tex.dtype == np.uint8:
                return tex
        except Exception:
            pass
    rng = np.random.default_rng(seed)
    acc = np.zeros((size, size, 3), dtype=np.float64)
    amp = 1.0
    total = 0.0
    for cells in (8, 16, 32, 64, 128):
        cells = min(cells, size)
        for ch in range(3):
            acc[:, :, ch] += amp * _value_noise(rng, size, cells)
        total += amp
        amp *= 0.6
    acc /= total
    img = np.clip(64.0 + 128.0 * acc, 0, 255).astype(np.uint8)
    smax = 160.0 * min(1.0, size / 512.0) if size < 512 else 160.0
    for _ in range(n_rects):
        cx, cy = rng.random(2) * size
        # log-uniform sizes favour many small, a few large rectangles
        w = float(np.exp(rng.uniform(np.log(10.0), np.log(smax))))
        h = float(np.exp(rng.uniform(np.log(10.0), np.log(smax))))
        ang = float(rng.uniform(0.0, np.pi))
        color = rng.integers(0, 256, size=3).astype(np.uint8)
        _fill_rot_rect(img, cx, cy, w, h, ang, color)
    img = np.ascontiguousarray(img)
    if cache:
        try:
            tmp = cache_path + f".{os.getpid()}.tmp.npy"
            np.save(tmp, img)
            os.replace(tmp, cache_path)
        except Exception:
            pass
    return img


def camera_path(n_frames: int, seed: int = PATH_SEED, drift: float = 0.0015) -> np.ndarray:
    """Scripted pose per frame: columns (x, y, z, pan, tilt, roll) in the units
    of CameraParams (/root/reference/include/camera_engine.hpp:44-74).

    SURVEY.md §8d: x = 0.5+0.0015 i+N(0,0.004), y = -0.3+0.05 sin(i/40)+N(0,0.004),
    roll = 180+3 sin(i/55)+N(0,0.35 deg), z = 0.7, pan 0, tilt 180.

    `drift` is the x velocity in world units per frame (0.0015 ~ 2.1 px/frame at 720p).  The
    full-lock benchmark uses drift = 0 (a hand-held camera over a fixed scene): with a steady
    drift a locked view leaves the anchor frame after a few hundred frames and the output
    degenerates to the border colour.
    """
    rng = np.random.default_rng(seed)
    i = np.arange(n_frames, dtype=np.float64)
    jx = rng.normal(0.0, 0.004, n_frames)
    jy = rng.normal(0.0, 0.004, n_frames)
    jr = rng.normal(0.0, 0.35, n_frames)
    out = np.empty((n_frames, 6), dtype=np.float64)
    out[:, 0] = 0.5 + drift * i + jx
    out[:, 1] = -0.3 + 0.05 * np.sin(i / 40.0) + jy
    out[:, 2] = 0.7
    out[:, 3] = 0.0
    out[:, 4] = 180.0
    out[:, 5] = 180.0 + 3.0 * np.sin(i / 55.0) + jr
    return out


def focal_for_width(width: int) -> float:
    """f = 1000 * W / 1280 keeps the field of view of main.cpp:29-36."""
    return 1000.0 * width / 1280.0
