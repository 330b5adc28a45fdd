"""Parametric spinning-LiDAR sweep generator (SURVEY.md §8d "Synthetic inputs").

There is no nuScenes/Waymo data in scope, so every test and benchmark draws
frames from this generator.  It is deliberately plain numpy on the host: the
product path starts where the reference's does, at a list of ``[N_i, 5]``
fp32 point arrays ``(x, y, z, intensity, dt)`` already range-filtered the way
``PointsRangeFilter`` does (reference ``mmdet3d/core/points/base_points.py:223-228``
keeps ``min < p < max`` strictly).

Presets
-------
``nuscenes``  32 beams, -30.67..+10.67 deg, 1090 azimuth steps, sensor 1.84 m
              above ground, range [-51.2,-51.2,-5, 51.2,51.2,3]
``waymo``     64 beams, -17.6..+2.4 deg, 2650 azimuth steps, range
              [-74.88,-74.88,-2, 74.88,74.88,4]
"""
from __future__ import annotations

import numpy as np

PRESETS = {
    "nuscenes": dict(beams=32, elev_deg=(-30.67, 10.67), az_steps=1090,
                     sensor_h=1.84,
                     pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)),
    "waymo": dict(beams=64, elev_deg=(-17.6, 2.4), az_steps=2650,
                  sensor_h=1.9,
                  pc_range=(-74.88, -74.88, -2.0, 74.88, 74.88, 4.0)),
}


def _one_sweep(rng, beams, elev_deg, az_steps, sensor_h, wall_r, wall_h, shift_x):
    elev = np.deg2rad(np.linspace(elev_deg[0], elev_deg[1], beams))
    az = np.linspace(-np.pi, np.pi, az_steps, endpoint=False)
    e, a = np.meshgrid(elev, az, indexing="ij")
    ce, se = np.cos(e), np.sin(e)
    # ground return: z = -sensor_h
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(se < 0, -sensor_h / se, np.inf)
    # one radial wall per azimuth degree (NaN radius = no wall there)
    sector = ((np.rad2deg(a) + 180.0) % 360.0).astype(np.int64) % 360
    r_w = wall_r[sector]
    h_w = wall_h[sector]
    t_wall = r_w / ce
    z_wall = t_wall * se
    hit_wall = np.isfinite(r_w) & (z_wall >= -sensor_h) & (z_wall <= h_w - sensor_h)
    t_wall = np.where(hit_wall, t_wall, np.inf)
    t = np.minimum(t_ground, t_wall)
    ok = np.isfinite(t) & (t < 120.0)
    t = t[ok] * (1.0 + 0.002 * rng.standard_normal(ok.sum()))
    x = t * ce[ok] * np.cos(a[ok]) + shift_x
    y = t * ce[ok] * np.sin(a[ok])
    z = t * se[ok]
    return x, y, z


def make_frame(seed: int, preset: str = "nuscenes", sweeps: int = 1,
               point_scale: float = 1.0) -> np.ndarray:
    """One range-filtered frame, ``[N, 5]`` float32, points shuffled.

    ``sweeps`` > 1 concatenates older sweeps with ``dt = 0.05*s`` and a
    0.5 m/sweep ego shift, as the multi-sweep loader does
    (reference ``mmdet3d/datasets/pipelines/loading.py:100-233``).
    ``point_scale`` multiplies the azimuth resolution (dense-grid stress).
    """
    p = PRESETS[preset]
    rng = np.random.default_rng(seed)
    wall_r = rng.uniform(5.0, 70.0, 360)
    wall_r[rng.random(360) < 0.3] = np.nan
    wall_h = rng.uniform(1.5, 6.0, 360)
    az_steps = int(round(p["az_steps"] * point_scale))
    chunks = []
    for s in range(sweeps):
        x, y, z = _one_sweep(rng, p["beams"], p["elev_deg"], az_steps,
                             p["sensor_h"], wall_r, wall_h, 0.5 * s)
        n = x.shape[0]
        chunks.append(np.stack([x, y, z, rng.uniform(0.0, 255.0, n),
                                np.full(n, 0.05 * s)], axis=1))
    pts = np.concatenate(chunks, axis=0).astype(np.float32)
    lo = np.asarray(p["pc_range"][:3], dtype=np.float32)
    hi = np.asarray(p["pc_range"][3:], dtype=np.float32)
    keep = np.all((pts[:, :3] > lo) & (pts[:, :3] < hi), axis=1)
    pts = pts[keep]
    rng.shuffle(pts, axis=0)
    return np.ascontiguousarray(pts)


def make_batch(seed: int, batch: int, **kw) -> list:
    """``batch`` frames with seeds ``seed*1000 + i`` (§8d: seed = 1000*rank + iter)."""
    return [make_frame(seed * 1000 + i, **kw) for i in range(batch)]
