"""Generates tests/golden/collision_golden.npz: decisions and distances of the rejectByCollisionOrNonTouching restatement
(oracle/hop_oracle_sdf.c, itself pinned against the reference's libigl by tests/test_sdf_oracle.py) on seeded grasp scenes.
Run from the repository root:  python tests/golden/make_collision_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
from hop_b200 import synth  # noqa: E402
from oracle import cpu_oracle as O  # noqa: E402

out = {}
for name, seed in (("ellipse", 21), ("cuboid", 22), ("tless", 23)):
    case = synth.make_collision_case(name, H=128, seed=seed)
    keep, reason, diag, amb = O.reject_by_collision(case, with_ambiguous=True)
    out[f"{name}_H"], out[f"{name}_seed"] = 128, seed
    out[f"{name}_reason"], out[f"{name}_diag"], out[f"{name}_ambiguous"] = reason, diag, amb
    print(name, np.bincount(reason, minlength=7), 'decisions hanging on a coin toss of the sign rules:', int(amb.sum()))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collision_golden.npz"), **out)
