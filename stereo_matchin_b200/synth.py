"""Deterministic synthetic stereo pairs for the named benchmark shapes.

Left image: clip(128 + 40 * (6 random-phase 2-D sinusoids, wavelengths 8-128 px) + N(0,12^2)
per pixel per channel).  Ground-truth disparity: 12 fronto-parallel rectangles over a slanted
background plane, integer d in [0, D-1].  Right image: R(x - d(x,y), y) = L(x,y), painted far to
near, holes filled from the left neighbour, + N(0,2^2) noise.  Returns RGBA8 (A = 255).
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (W, H, D, seed)            BASELINE.json configs
    "cfg1a_sukub": (384, 288, 16, 1),
    "cfg2_teddy_shape": (450, 375, 61, 2),
    "cfg3_1800x1500_d256": (1800, 1500, 256, 3),
    "cfg4_3840x2160_d256": (3840, 2160, 256, 4),
    "cfg5_1280x720_d128": (1280, 720, 128, 5000),
}


def make_pair(W: int, H: int, D: int, seed: int):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    tex = np.zeros((H, W), np.float32)
    for _ in range(6):
        lam = rng.uniform(8.0, 128.0)
        th = rng.uniform(0.0, 2 * np.pi)
        ph = rng.uniform(0.0, 2 * np.pi)
        tex += np.sin(2 * np.pi * (xx * np.cos(th) + yy * np.sin(th)) / lam + ph).astype(np.float32)
    left = 128.0 + 40.0 * tex[..., None] / 2.0 + rng.normal(0.0, 12.0, (H, W, 3)).astype(np.float32)
    left = np.clip(np.rint(left), 0, 255).astype(np.uint8)

    dmax = max(D - 1, 0)
    # slanted background plane
    disp = (0.15 * dmax + 0.25 * dmax * (xx / max(W - 1, 1)) + 0.10 * dmax * (yy / max(H - 1, 1)))
    disp = np.clip(np.rint(disp), 0, dmax).astype(np.int32)
    for _ in range(12):
        w, h = int(rng.integers(max(W // 12, 1), max(W // 3, 2))), int(rng.integers(max(H // 12, 1), max(H // 3, 2)))
        x0, y0 = int(rng.integers(0, max(W - w, 1))), int(rng.integers(0, max(H - h, 1)))
        d = int(rng.integers(0, dmax + 1))
        disp[y0:y0 + h, x0:x0 + w] = np.maximum(disp[y0:y0 + h, x0:x0 + w], d)

    right = np.zeros((H, W, 3), np.float32)
    filled = np.zeros((H, W), bool)
    rows = np.arange(H)[:, None].repeat(W, 1)
    # paint far (small d) to near (large d): nearer surfaces overwrite
    order = np.argsort(disp, axis=None, kind="stable")
    ys, xs = rows.reshape(-1)[order], np.tile(np.arange(W), H)[order]
    xr = xs - disp.reshape(-1)[order]
    ok = xr >= 0
    right[ys[ok], xr[ok]] = left[ys[ok], xs[ok]]
    filled[ys[ok], xr[ok]] = True
    # fill holes from the left neighbour (first column falls back to the left image)
    for x in range(W):
        hole = ~filled[:, x]
        if hole.any():
            right[hole, x] = right[hole, x - 1] if x > 0 else left[hole, 0]
    right = right + rng.normal(0.0, 2.0, (H, W, 3)).astype(np.float32)
    right = np.clip(np.rint(right), 0, 255).astype(np.uint8)

    a = np.full((H, W, 1), 255, np.uint8)
    return np.concatenate([left, a], 2), np.concatenate([right, a], 2), disp


def make_config(name: str, index: int = 0):
    W, H, D, seed = CONFIGS[name]
    L, R, gt = make_pair(W, H, D, seed + index)
    return L, R, gt, D
