"""The shift/mask path of the voxel kernels computes ONE low-scale coordinate per axis as floor((p - lo) * (1/v)) and
falls back to the IEEE divide only inside a band of width qeps around integers (csrc/voxel_geom.cuh::vox_coord_try,
VoxGeom::qeps).  This file replays that arithmetic in numpy float32 (IEEE, same operations, same constants) and checks
the claim the kernels rely on: outside the band the product's floor equals the floor of the reference's quotient
(voxelization_cpu.cpp:22-31) — on random points, on points a few ulps around every kind of voxel face, and far outside
the range — and that coarser scales are exact shifts of the low-scale coordinate."""
import numpy as np
import pytest

from oracle import geomae_oracle as O
from oracle.make_golden import DENSE_GEOMETRY, WAYMO_GEOMETRY

GEOMETRIES = {"nuscenes": {}, "waymo": WAYMO_GEOMETRY, "dense": DENSE_GEOMETRY}
F = np.float32


def host_constants(cfg, axis):
    """gm_make_geom: grid = ceil of the fp32 quotient, rvs = 1.0f / v, qeps = (grid + 2) * 4.8e-7f + 1e-6f."""
    lo, hi = F(cfg.pc_range[axis]), F(cfg.pc_range[axis + 3])
    v = F(cfg.sub_voxel_size_low[axis])
    grid = int(np.ceil(F(F(hi - lo) / v)))
    rvs = F(F(1.0) / v)
    qeps = F(F(F(grid + 2) * F(4.8e-7)) + F(1e-6))
    return lo, v, grid, rvs, qeps


def sample_points(rng, lo, v, grid, n):
    span = F(v) * F(grid)
    uniform = rng.uniform(float(lo) - 0.3 * float(span), float(lo) + 1.3 * float(span), n).astype(F)
    # voxel faces of the low scale (hence of every scale) and a few ulps either side
    k = rng.integers(-4, grid + 5, n)
    face = (k.astype(F) * F(v) + F(lo)).astype(F)
    for _ in range(3):
        step = rng.integers(-1, 2, n)
        face = np.where(step > 0, np.nextafter(face, F(1e30)), np.where(step < 0, np.nextafter(face, F(-1e30)), face))
    far = rng.uniform(-1e6, 1e6, n // 8).astype(F)
    return np.concatenate([uniform, face, far, np.array([lo, lo + span, 0.0], F)])


@pytest.mark.parametrize("name", sorted(GEOMETRIES))
def test_reciprocal_multiply_coordinate_equals_ieee_divide_outside_the_band(name):
    cfg = O.PathConfig(**GEOMETRIES[name])
    rng = np.random.default_rng(17)
    for axis in range(3):
        lo, v, grid, rvs, qeps = host_constants(cfg, axis)
        assert qeps < 0.25
        p = sample_points(rng, lo, v, grid, 2_000_000)
        d = (p - lo).astype(F)                                  # __fsub_rn
        q = (d * rvs).astype(F)                                 # __fmul_rn
        f = np.floor(q)
        fr = (q - f).astype(F)
        redo = ~((fr > qeps) & (fr < F(1.0) - qeps))
        exact = np.floor((d / v).astype(F))                     # floorf(__fdiv_rn(d, v))
        clamp = lambda c: np.clip(c, 0, grid - 1)               # noqa: E731
        fast = clamp(f[~redo].astype(np.int64))
        ref = clamp(exact[~redo].astype(np.int64))
        assert np.array_equal(fast, ref), (name, axis, int((fast != ref).sum()))
        # the band must stay a rare slow path for in-range points
        inside = (p >= lo) & (p < lo + F(v) * F(grid))
        uniform_part = slice(0, 2_000_000)
        frac = redo[uniform_part][inside[uniform_part]].mean()
        assert frac < 0.02, (name, axis, frac)


@pytest.mark.parametrize("name", sorted(GEOMETRIES))
def test_coarser_scales_are_shifts_of_the_low_scale_coordinate(name):
    """v_s = v_low * 2^k and grid_low = grid_s * 2^k make the independent IEEE coordinate of scale s equal
    to the low-scale coordinate >> k (dividing by v * 2^k only changes the quotient's exponent)."""
    cfg = O.PathConfig(**GEOMETRIES[name])
    rng = np.random.default_rng(23)
    n = 400_000
    lo3, hi3 = np.array(cfg.pc_range[:3], F), np.array(cfg.pc_range[3:], F)
    pts = np.concatenate([rng.uniform(lo3 - 5, hi3 + 5, (n, 3)).astype(F), np.zeros((n, 2), F)], axis=1)
    low = O.dynamic_voxelize(pts, cfg.sub_voxel_size_low, cfg.pc_range)       # (z, y, x)
    for size in (cfg.sub_voxel_size_med, cfg.voxel_size):
        coarse = O.dynamic_voxelize(pts, size, cfg.pc_range)
        for col, axis in ((0, 2), (1, 1), (2, 0)):
            ratio = F(size[axis]) / F(cfg.sub_voxel_size_low[axis])
            k = int(round(np.log2(float(ratio))))
            assert F(cfg.sub_voxel_size_low[axis]) * F(2 ** k) == F(size[axis]), (name, size, axis)
            assert np.array_equal(coarse[:, col], low[:, col] >> k), (name, size, axis)
