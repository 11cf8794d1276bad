"""Seeded synthetic stereo pairs.

There are no datasets in this environment, so every test and the bench use images made here:
  left  = 1/f-ish noise (3 octaves of bilinearly up-sampled uniform noise)
          + a few hundred/thousand random rectangles and checker patches with contrast >= 40
            (guarantees FAST corners in most NMS cells), clipped to u8
  right = left shifted per row band by a disparity d in [4, 64] px + independent N(0, 2) noise
so that left<->right matches exist.  Pure NumPy, deterministic for a given (seed, height, width).
"""
from __future__ import annotations

import math

import numpy as np


def _upsample_bilinear(a: np.ndarray, h: int, w: int) -> np.ndarray:
    sh, sw = a.shape
    ys = np.linspace(0, sh - 1, h)
    xs = np.linspace(0, sw - 1, w)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, sh - 1)
    x1 = np.minimum(x0 + 1, sw - 1)
    wy = (ys - y0)[:, None]
    wx = (xs - x0)[None, :]
    top = a[y0][:, x0] * (1 - wx) + a[y0][:, x1] * wx
    bot = a[y1][:, x0] * (1 - wx) + a[y1][:, x1] * wx
    return top * (1 - wy) + bot * wy


def textured_image(height: int, width: int, seed: int, n_shapes: int | None = None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 110.0)
    for octave, amp in ((16, 40.0), (8, 25.0), (3, 12.0)):
        coarse = rng.uniform(-1.0, 1.0, size=(max(2, height // octave), max(2, width // octave)))
        img += amp * _upsample_bilinear(coarse, height, width)
    if n_shapes is None:
        n_shapes = max(400, (height * width) // 450)
    for _ in range(n_shapes):
        h = int(rng.integers(4, 28))
        w = int(rng.integers(4, 28))
        y = int(rng.integers(0, max(1, height - h)))
        x = int(rng.integers(0, max(1, width - w)))
        contrast = float(rng.integers(40, 110)) * (1 if rng.random() < 0.5 else -1)
        if rng.random() < 0.25:  # checker patch
            c = int(rng.integers(2, 6))
            yy, xx = np.mgrid[0:h, 0:w]
            img[y:y + h, x:x + w] += contrast * (((yy // c) + (xx // c)) % 2)
        else:
            img[y:y + h, x:x + w] += contrast
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def stereo_pair(height: int, width: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    """Return (left, right) u8 images, C-contiguous, shape (height, width)."""
    left = textured_image(height, width, seed)
    rng = np.random.default_rng(seed + 1_000_003)
    right = np.empty_like(left)
    band = 32
    for y0 in range(0, height, band):
        d = int(rng.integers(4, 65))
        rows = left[y0:y0 + band]
        shifted = np.empty_like(rows)
        shifted[:, :width - d] = rows[:, d:]       # a point at uL appears at uR = uL - d
        shifted[:, width - d:] = rows[:, width - 1:width]
        right[y0:y0 + band] = shifted
    noisy = right.astype(np.float64) + rng.normal(0.0, 2.0, size=right.shape)
    right = np.clip(np.rint(noisy), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


def degenerate_images(height: int, width: int) -> dict[str, np.ndarray]:
    """Edge cases used by the parity tests (tie-breaks, empty outputs)."""
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:height, 0:width]
    out = {
        "zeros": np.zeros((height, width), np.uint8),
        "full": np.full((height, width), 255, np.uint8),
        "noise": rng.integers(0, 256, size=(height, width), dtype=np.uint8),
        "vstep": np.where(xx < width // 2, 30, 220).astype(np.uint8),
        "hstep": np.where(yy < height // 2, 30, 220).astype(np.uint8),
        # 8x8 checkerboard: every corner has the same score -> exercises every tie-break rule
        "checker": (((yy // 8) + (xx // 8)) % 2 * 200 + 20).astype(np.uint8),
    }
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


# ---- a synthetic SearchByProjection problem (KITTI-like intrinsics) ----
SBP_K = dict(fx=718.856, fy=718.856, cx=607.19, cy=185.2)
SBP_BOUNDS = dict(min_x=0.0, max_x=1241.0, min_y=0.0, max_y=376.0)
SBP_MBF = 386.1448


def projection_scene(n_cur=1500, n_last=1200, seed=0, noise_px=3.0, clustered=False, dup_desc=False):
    """A current frame (keypoints on a pixel grid, random descriptors) and last-frame map points that project near a
    subset of them under a small camera motion; descriptors differ by a few bits so that matches exist."""
    rng = np.random.default_rng(seed)
    if clustered:   # everything inside two grid cells: long candidate lists, many ties
        x = rng.integers(300, 330, size=n_cur).astype(np.float32)
        y = rng.integers(100, 112, size=n_cur).astype(np.float32)
    else:
        x = rng.integers(5, 1236, size=n_cur).astype(np.float32)
        y = rng.integers(5, 371, size=n_cur).astype(np.float32)
    octave = rng.integers(0, 8, size=n_cur).astype(np.int32)
    angle = rng.uniform(0, 360, size=n_cur).astype(np.float32)
    desc = rng.integers(0, 256, size=(n_cur, 32), dtype=np.uint8)
    if dup_desc:
        desc[:] = desc[0]
    uright = np.where(rng.random(n_cur) < 0.6, x - rng.uniform(2, 60, size=n_cur), -1.0).astype(np.float32)
    occupied = (rng.random(n_cur) < 0.15).astype(np.uint8)
    cur = dict(x=x, y=y, octave=octave, angle=angle, uright=uright, occupied=occupied, desc=desc)
    # camera pose of the current frame
    a = 0.02
    R = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]], np.float32)
    t = np.array([0.05, -0.02, 0.3], np.float32)
    src = rng.integers(0, n_cur, size=n_last)
    z = np.where(uright[src] > 0, SBP_MBF / np.maximum(x[src] - uright[src], 0.5), rng.uniform(4, 60, size=n_last)).astype(np.float64)
    z[rng.random(n_last) < 0.03] *= -1                      # some behind the camera
    u = x[src] + rng.normal(0, noise_px, size=n_last)
    v = y[src] + rng.normal(0, noise_px, size=n_last)
    Pc = np.stack([(u - SBP_K["cx"]) * z / SBP_K["fx"], (v - SBP_K["cy"]) * z / SBP_K["fy"], z])
    Pw = (R.astype(np.float64).T @ (Pc - t[:, None].astype(np.float64))).astype(np.float32)
    ld = desc[src].copy()
    flips = rng.integers(0, 256, size=(n_last, 6))
    for i in range(n_last):
        for b in flips[i, : rng.integers(0, 7)]:
            ld[i, b >> 3] ^= np.uint8(1 << (b & 7))
    strangers = rng.random(n_last) < 0.1
    ld[strangers] = rng.integers(0, 256, size=(int(strangers.sum()), 32), dtype=np.uint8)
    la = (angle[src] + np.where(rng.random(n_last) < 0.8, rng.normal(5, 3, size=n_last), rng.uniform(-180, 180, size=n_last))).astype(np.float32) % np.float32(360)
    lo = np.clip(octave[src] + rng.integers(-1, 2, size=n_last), 0, 7).astype(np.int32)
    last = dict(P=Pw, octave=lo, angle=la.astype(np.float32), desc=ld)
    return last, cur, R.ravel().copy(), t


