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
