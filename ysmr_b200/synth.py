"""Seeded synthetic rod/coccoid bacteria video generator.

The same bytes are fed to the CPU oracle and to the CUDA path.  Two back ends share one scene
description (positions / orientations per frame, produced by :func:`make_scene` with numpy):

* :func:`render_frames` -- numpy, used by tests, golden fixtures and small CPU runs;
* :func:`render_frames_torch` -- torch (any device), used by ``bench.py`` to fill HBM with thousands of
  frames quickly.  Noise comes from a seeded ``torch.Generator`` so it is reproducible per device type,
  and the CPU baseline of the bench copies a prefix of *these* bytes back to the host.

Scene model: constant background + N(0, sigma) noise, anti-aliased filled ellipses (rods or discs) that
move with an AR(1) velocity and a random-walk orientation; frames are grey, expanded to BGR (B=G=R) the
way ``cv2.VideoCapture.read`` delivers a grey file to the reference loop (track_eval.py:159).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

__all__ = ['Scene', 'SceneConfig', 'make_scene', 'render_frames', 'render_frames_torch', 'to_bgr', 'CONFIGS']

PATCH = 17  # rendering patch edge (px); semi-axes up to ~7 px fit
SS = 4      # supersampling per axis for the anti-aliased coverage


@dataclass(frozen=True)
class SceneConfig:
    width: int = 1228
    height: int = 922
    n_frames: int = 300
    n_cells: int = 50
    seed: int = 0
    background: float = 60.0
    intensity: float = 150.0
    noise_sigma: float = 3.0
    semi_major: float = 4.0
    semi_minor: float = 1.3
    margin: float = 100.0
    fps: float = 30.0


# The five BASELINE.json configurations (frames count is what the config names; tests use prefixes).
CONFIGS = {
    'cfg1': SceneConfig(n_frames=300),
    'cfg2': SceneConfig(n_frames=9000),
    'cfg3': SceneConfig(n_frames=9000, n_cells=2000, margin=20.0),
    'cfg4': SceneConfig(width=2048, height=2048, n_frames=3000, n_cells=200, background=160.0,
                        intensity=80.0, noise_sigma=1.5, semi_major=2.5, semi_minor=2.5),
    'cfg5': SceneConfig(n_frames=54000),
}


@dataclass
class Scene:
    cfg: SceneConfig
    x: np.ndarray      # (n_frames, n_cells) float64 centre x
    y: np.ndarray      # (n_frames, n_cells) float64 centre y
    theta: np.ndarray  # (n_frames, n_cells) float64 orientation, radians


def make_scene(cfg: SceneConfig) -> Scene:
    """Trajectories: v_t = 0.9 v_{t-1} + N(0, 0.3); theta_t = theta_{t-1} + N(0, 2 deg); reflecting walls."""
    rng = np.random.default_rng(cfg.seed)
    n, f = cfg.n_cells, cfg.n_frames
    x = np.empty((f, n)); y = np.empty((f, n)); th = np.empty((f, n))
    px = rng.uniform(cfg.margin, cfg.width - cfg.margin, n)
    py = rng.uniform(cfg.margin, cfg.height - cfg.margin, n)
    vx = rng.normal(0, 0.3, n); vy = rng.normal(0, 0.3, n)
    a = rng.uniform(0, np.pi, n)
    lo = 12.0
    for t in range(f):
        x[t], y[t], th[t] = px, py, a
        vx = 0.9 * vx + rng.normal(0, 0.3, n)
        vy = 0.9 * vy + rng.normal(0, 0.3, n)
        px = px + vx; py = py + vy
        # reflect at a 12 px inner wall so patches never leave the frame
        bx = (px < lo) | (px > cfg.width - 1 - lo); by = (py < lo) | (py > cfg.height - 1 - lo)
        vx = np.where(bx, -vx, vx); vy = np.where(by, -vy, vy)
        px = np.clip(px, lo, cfg.width - 1 - lo); py = np.clip(py, lo, cfg.height - 1 - lo)
        a = a + rng.normal(0, np.deg2rad(2.0), n)
    return Scene(cfg, x, y, th)


def _coverage_np(cx, cy, theta, a, b):
    """Anti-aliased ellipse coverage on the PATCH x PATCH window around round(cx), round(cy).

    Returns (alpha[n, PATCH, PATCH], x0[n], y0[n]) with x0/y0 the window origin (integer pixel)."""
    ix = np.rint(cx).astype(np.int64); iy = np.rint(cy).astype(np.int64)
    x0 = ix - PATCH // 2; y0 = iy - PATCH // 2
    sub = (np.arange(SS) + 0.5) / SS - 0.5
    gx = np.arange(PATCH)[None, :, None] + sub[None, None, :]            # (1, PATCH, SS)
    gx = gx.reshape(1, PATCH * SS)
    dx = (x0[:, None] + gx) - cx[:, None]                                 # (n, PATCH*SS)
    dy = (y0[:, None] + gx) - cy[:, None]
    c = np.cos(theta)[:, None, None]; s = np.sin(theta)[:, None, None]
    DX = dx[:, None, :]; DY = dy[:, :, None]                              # (n, ys, xs)
    u = DX * c + DY * s
    v = -DX * s + DY * c
    inside = ((u / a) ** 2 + (v / b) ** 2) <= 1.0
    n = cx.shape[0]
    alpha = inside.reshape(n, PATCH, SS, PATCH, SS).mean(axis=(2, 4))
    return alpha, x0, y0


def render_frames(scene: Scene, start: int = 0, stop: int | None = None) -> np.ndarray:
    """Grey frames [start, stop) as uint8 (n, H, W).  Noise stream is keyed by (seed, frame index) so any
    frame range renders identically regardless of how the video is chunked or sharded."""
    cfg = scene.cfg
    stop = cfg.n_frames if stop is None else stop
    out = np.empty((stop - start, cfg.height, cfg.width), np.uint8)
    for k, t in enumerate(range(start, stop)):
        rng = np.random.default_rng([cfg.seed, 7919, t])
        img = np.full((cfg.height, cfg.width), cfg.background, np.float32)
        alpha, x0, y0 = _coverage_np(scene.x[t], scene.y[t], scene.theta[t], cfg.semi_major, cfg.semi_minor)
        amp = np.float32(cfg.intensity - cfg.background)
        for i in range(alpha.shape[0]):
            ys, xs = int(y0[i]), int(x0[i])
            win = img[ys:ys + PATCH, xs:xs + PATCH]
            # overlapping cells: max blend (a cell does not get brighter where two overlap)
            np.maximum(win, cfg.background + amp * alpha[i].astype(np.float32), out=win) if amp > 0 else \
                np.minimum(win, cfg.background + amp * alpha[i].astype(np.float32), out=win)
        img += rng.normal(0.0, cfg.noise_sigma, img.shape).astype(np.float32)
        out[k] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def to_bgr(grey: np.ndarray) -> np.ndarray:
    """(n, H, W) grey -> (n, H, W, 3) BGR with B=G=R, C-contiguous, as cap.read() yields for a grey file."""
    return np.ascontiguousarray(np.repeat(grey[..., None], 3, axis=-1))


def render_frames_torch(scene: Scene, start: int, stop: int, device, channels: int = 1, out=None):
    """Torch renderer (GPU capable).  Returns uint8 (n, H, W) or (n, H, W, 3).  Same scene geometry as the
    numpy renderer; the noise stream differs (torch generator), which is fine because every consumer of
    these frames (CUDA path, CPU baseline in bench.py) reads the same tensor."""
    import torch
    cfg = scene.cfg
    n = stop - start
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    img = torch.full((n, cfg.height, cfg.width), float(cfg.background), device=dev, dtype=torch.float32)
    alpha, x0, y0 = [], [], []
    for t in range(start, stop):
        a_, x_, y_ = _coverage_np(scene.x[t], scene.y[t], scene.theta[t], cfg.semi_major, cfg.semi_minor)
        alpha.append(a_); x0.append(x_); y0.append(y_)
    alpha = torch.from_numpy(np.stack(alpha).astype(np.float32)).to(dev)          # (n, cells, P, P)
    x0 = torch.from_numpy(np.stack(x0)).to(dev); y0 = torch.from_numpy(np.stack(y0)).to(dev)
    amp = float(cfg.intensity - cfg.background)
    ar = torch.arange(PATCH, device=dev)
    yy = (y0[:, :, None, None] + ar[None, None, :, None]).expand(-1, -1, PATCH, PATCH)
    xx = (x0[:, :, None, None] + ar[None, None, None, :]).expand(-1, -1, PATCH, PATCH)
    ff = torch.arange(n, device=dev)[:, None, None, None].expand_as(yy)
    flat = (ff * cfg.height + yy) * cfg.width + xx
    val = cfg.background + amp * alpha
    img.view(-1).scatter_reduce_(0, flat.reshape(-1), val.reshape(-1), reduce='amax' if amp > 0 else 'amin')
    for i in range(n):                      # one noise stream per FRAME, so the bytes do not depend on how the caller batches
        g.manual_seed(cfg.seed * 1000003 + start + i)
        img[i] += torch.randn(img.shape[1:], generator=g, device=dev, dtype=torch.float32) * cfg.noise_sigma
    grey = img.round_().clamp_(0, 255).to(torch.uint8)
    if channels == 1:
        if out is not None:
            out.copy_(grey); return out
        return grey
    if out is None:
        out = torch.empty((n, cfg.height, cfg.width, 3), device=dev, dtype=torch.uint8)
    out.copy_(grey[..., None].expand(-1, -1, -1, 3))
    return out
