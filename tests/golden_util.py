"""Shared helpers: load a golden fixture and rebuild its inputs without /root/reference."""
import os
import zlib

import numpy as np

from oracle.make_golden import CASES, GOLDEN_DIR, case_cfg, case_frames  # noqa: F401  (pure-python, no reference needed)


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    case = CASES[name]
    frames = case_frames(case)
    crc = np.array([zlib.crc32(f.tobytes()) for f in frames], np.int64)
    assert (crc == g["points_crc"]).all(), "synthetic generator drifted from the committed golden inputs"
    return case, case_cfg(case), frames, g


def align_sign(ref, mine):
    """The reference's normal sign is a LAPACK artefact: flip ours onto it."""
    s = np.sign((ref * mine).sum(-1, keepdims=True))
    s[s == 0] = 1
    return mine * s


def assert_same_step(a, b, tol=2e-5, curv_tol=1e-3):
    """Two runs of the SAME training step (same weights, frames, mask split), given as loss dicts.  Five of the six
    terms repeat to float-atomics noise (~1e-7).  `loss_curv_around` regresses unit normals, and the normal of a
    pillar with a near-degenerate scatter matrix flips with the last bits of its sub-voxel centroids (which come
    from order-dependent float atomics): repeated identical steps move that term in quanta of 2e-5..1.2e-4
    (tools/repeat_forward.py), so it gets its own, looser bound."""
    assert set(a) == set(b)
    for k in a:
        x, y = float(a[k]), float(b[k])
        bound = curv_tol if k == "loss_curv_around" else tol
        assert abs(x - y) <= bound * abs(y), (k, x, y)
