"""Generates tests/golden/icp_lcp_small.npz: seeded inputs and the oracle's outputs for them.

Run from the repo root:  python tests/golden/make_golden.py
The ICP outputs are produced with the LM step routed through oracle/_ref (the reference tree's own Eigen
Levenberg-Marquardt) when it is built, i.e. they are outputs of reference code wherever reference code exists for
this path; the C restatement must reproduce them (tests/test_oracle.py::test_golden_fixtures).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
from hop_b200 import synth  # noqa: E402
from oracle import cpu_oracle as O  # noqa: E402

m, mn = synth.make_model("ellipse", 1500, seed=1)
s, sn, conf, gt = synth.make_scene("ellipse", 300, seed=2)
hyp = synth.make_hypotheses(gt, 16, seed=3)
used_ref = O.ref() is not None
if used_ref:
    O.use_ref_lm(True)
refined, iters, conv = O.refine_by_icp(s, sn, m, mn, hyp, nthreads=1)
O.use_ref_lm(False) if used_ref else None
_, scores = O.select_best(s, sn, m, mn, refined, nthreads=1)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icp_lcp_small.npz")
np.savez_compressed(out, m=m, mn=mn, s=s, sn=sn, conf=conf, gt=gt, hyp=hyp, refined=refined, iters=iters, conv=conv,
                    scores=scores, lm_backend=np.array("reference-eigen" if used_ref else "c-port"))
print("wrote", out, "lm backend:", "reference Eigen (oracle/_ref)" if used_ref else "C port")
