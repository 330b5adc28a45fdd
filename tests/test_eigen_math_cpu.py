"""float32 replay of k_geom's eigen-solve (csrc/geom_targets.cu: cyclic Jacobi in fp32, eigenvalues as fp64 Rayleigh
quotients of the fp32 moment matrix) on the moment matrices of a real synthetic frame, against LAPACK in fp64.
Pins the algorithm's accuracy without a GPU: the bounds asserted here are the ones the GPU parity tests rely on.
(The kernel forms the rotation angle with reciprocal-unit divides; the angle only has to be approximately the
annihilating one, c and s are formed from it the same way as here, so the replay uses plain fp32 division.)"""
import numpy as np

from geomae_b200.synthetic import make_frame
from oracle import geomae_oracle as O

F = np.float32


def jacobi3_f32(a):
    """a: [n,3,3] float32 symmetric.  Vectorised replay of jacobi3f: same rotation order (0,1), (0,2), (1,2), same
    formulas, per-matrix convergence test 'off <= 1e-15 * diag' evaluated in fp32, at most 8 sweeps."""
    a = a.astype(F).copy()
    n = a.shape[0]
    v = np.tile(np.eye(3, dtype=F), (n, 1, 1))
    active = np.ones(n, bool)
    for _ in range(8):
        off = a[:, 0, 1] ** 2 + a[:, 0, 2] ** 2 + a[:, 1, 2] ** 2
        diag = a[:, 0, 0] ** 2 + a[:, 1, 1] ** 2 + a[:, 2, 2] ** 2
        active &= ~((off <= F(1e-15) * diag) | (off == 0))
        if not active.any():
            break
        for p, q in ((0, 1), (0, 2), (1, 2)):
            r = 3 - p - q
            apq = a[:, p, q]
            rot = active & (np.abs(apq) >= F(1e-30))
            small = active & ~rot
            a[small, p, q] = a[small, q, p] = 0
            if not rot.any():
                continue
            i = np.where(rot)[0]
            apq = a[i, p, q]
            with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
                theta = ((a[i, q, q] - a[i, p, p]) / (F(2) * apq)).astype(F)
                t = (np.sign(theta) + (theta == 0)) / (np.abs(theta) + np.sqrt(theta * theta + F(1)))
            t = np.nan_to_num(t.astype(F), nan=0.0)
            c = (F(1) / np.sqrt(t * t + F(1))).astype(F)
            s = (t * c).astype(F)
            arp, arq = a[i, r, p].copy(), a[i, r, q].copy()
            a[i, p, p] = a[i, p, p] - t * apq
            a[i, q, q] = a[i, q, q] + t * apq
            a[i, p, q] = a[i, q, p] = 0
            a[i, r, p] = a[i, p, r] = c * arp - s * arq
            a[i, r, q] = a[i, q, r] = s * arp + c * arq
            vp, vq = v[i, :, p].copy(), v[i, :, q].copy()
            v[i, :, p] = c[:, None] * vp - s[:, None] * vq
            v[i, :, q] = s[:, None] * vp + c[:, None] * vq
    return v


def test_fp32_jacobi_with_fp64_rayleigh_quotients_matches_lapack():
    cfg = O.PathConfig()
    frames = [make_frame(61), make_frame(62, sweeps=2)]
    tgt = O.geometric_targets(frames, cfg, np.arange(4))
    cov = tgt["cov"].astype(F)                                   # [V,3,3] fp32 moments, (z,y,x)
    v = jacobi3_f32(cov)
    a64 = cov.astype(np.float64)
    e = v.astype(np.float64)
    lam = np.abs(np.einsum("nij,nik,nkj->nj", e, a64, e) / np.einsum("nij,nij->nj", e, e))   # Rayleigh quotients
    order = np.argsort(-lam, axis=1, kind="stable")
    lam = np.take_along_axis(lam, order, axis=1)
    # fp64 LAPACK; the kernel reports |lambda| (the fp32-rounded Gram matrix can be indefinite by an ulp), descending
    w = -np.sort(-np.abs(np.linalg.eigvalsh(a64)), axis=1)
    s0 = np.maximum(w[:, 0], 1e-30)
    # eigenvalues: second order in the fp32 eigenvector error -> far below fp32 resolution of the matrix itself
    assert (np.abs(lam - w) <= 1e-9 * s0[:, None] + 1e-30).all(), float((np.abs(lam - w) / s0[:, None]).max())
    # normal = eigenvector of the smallest eigenvalue: residual on all pillars, direction where well conditioned
    nrm = np.take_along_axis(e, order[:, None, 2:3].repeat(3, 1), axis=2)[:, :, 0]
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    res = np.einsum("nij,nj->ni", a64, nrm) - lam[:, 2:3] * nrm
    assert (np.abs(res).max(axis=1) <= 2e-6 * s0 + 1e-12).all(), float((np.abs(res).max(axis=1) / s0).max())
    wv, vec = np.linalg.eigh(a64)
    ref = np.take_along_axis(vec, np.argmin(np.abs(wv), axis=1)[:, None, None].repeat(3, 1), axis=2)[:, :, 0]
    well = (w[:, 1] - w[:, 2]) > 1e-3 * s0
    assert well.mean() > 0.3
    ang = 1.0 - np.abs(np.einsum("ni,ni->n", nrm, ref))
    assert ang[well].max() < 1e-6, float(ang[well].max())       # < 1.5e-3 rad
