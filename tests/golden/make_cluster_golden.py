"""Generates tests/golden/cluster_poses.npz: seeded hypothesis sets and the clusters the restatement of clusterPoses built
on the reference tree's own Eigen (oracle/_ref/ref_cluster.cpp) keeps for them.  Run from the repo root:
    python tests/golden/make_cluster_golden.py        (needs oracle/_ref, i.e. /root/reference at build time)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "icra20-hand-object-pose_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oracle import cpu_oracle as O  # noqa: E402
import test_cluster as T  # noqa: E402

out = {"n_cases": len(T.CASES)}
for k, c in enumerate(T.CASES):
    hyp, scores = T._hypotheses(c["seed"], c["n"], *c.get("spread", (25.0, 0.02)))
    out[f"hyp{k}"], out[f"scores{k}"], out[f"angle{k}"], out[f"dist{k}"], out[f"sym{k}"] = hyp, scores, c["angle"], c["dist"], np.array(c["sym"])
    out[f"keep{k}"] = O.ref_cluster_poses(hyp, scores, c["angle"], c["dist"], c["sym"])
    print("case", k, "kept", len(out[f"keep{k}"]), "of", c["n"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cluster_poses.npz"), **out)
