"""clusterPoses (host function of libhop) against the restatement built on the reference tree's own Eigen (oracle/_ref) and
against a committed golden fixture generated from it."""
import os

import numpy as np
import pytest

from hop_b200 import capi, synth
from oracle import cpu_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cluster_poses.npz")
has_ref = O.ref() is not None and hasattr(O.ref(), "hop_ref_cluster_poses")


def _hypotheses(seed, n, spread_deg=25.0, spread_t=0.02):
    rng = np.random.default_rng(seed)
    gt = synth.make_gt_pose(rng).astype(np.float32)
    hyp = synth.make_hypotheses(gt, n, seed=seed + 1, rot_sigma_deg=spread_deg, trans_sigma=spread_t, random_frac=0.2)
    scores = (rng.integers(0, 40, n) / 100.0).astype(np.float32)   # LCP = k / 100: many ties, broken by id
    return hyp, scores


CASES = [dict(seed=1, n=600, angle=30.0, dist=0.015, sym=(360.0, 360.0, 360.0)),      # main_realdata_auto.cpp:198
         dict(seed=2, n=600, angle=5.0, dist=0.003, sym=(360.0, 360.0, 360.0), spread=(3.0, 0.002)),   # :200
         dict(seed=3, n=400, angle=30.0, dist=0.015, sym=(180.0, 180.0, 0.0)),        # a free z axis + two-fold x / y
         dict(seed=4, n=400, angle=10.0, dist=0.05, sym=(-1.0, 90.0, 360.0))]


@pytest.mark.skipif(not has_ref, reason="oracle/_ref not built")
@pytest.mark.parametrize("case", CASES)
def test_cluster_matches_reference_eigen_restatement(case):
    hyp, scores = _hypotheses(case["seed"], case["n"], *case.get("spread", (25.0, 0.02)))
    got = capi.cluster_poses(hyp, scores, case["angle"], case["dist"], case["sym"])
    ref = O.ref_cluster_poses(hyp, scores, case["angle"], case["dist"], case["sym"])
    assert np.array_equal(got, ref)
    assert 1 < len(got) < case["n"]
    # cluster order = (score desc, id asc) order of the survivors
    assert np.array_equal(got, sorted(got, key=lambda k: (-scores[k], k)))
    # explicit ids are honoured in the tie-break
    ids = np.arange(case["n"])[::-1].astype(np.int32)
    assert np.array_equal(capi.cluster_poses(hyp, scores, case["angle"], case["dist"], case["sym"], ids),
                          O.ref_cluster_poses(hyp, scores, case["angle"], case["dist"], case["sym"], ids))


@pytest.mark.skipif(not has_ref, reason="oracle/_ref not built")
def test_euler_angles_match_eigen():
    rng = np.random.default_rng(0)
    for _ in range(300):
        R = synth.random_rotation(rng)
        P = np.eye(4, dtype=np.float32); P[:3, :3] = R
        # two poses closer than any threshold in translation, euler-compared only: equality of the decision is what matters,
        # so compare through the clustering itself with a tight angle and no symmetry
        Q = P.copy(); Q[:3, :3] = (R @ synth._rot_from_rotvec(rng.normal(0, 0.02, 3))).astype(np.float32)
        both = np.stack([P, Q])
        for ang in (0.5, 1.0, 2.0):
            assert np.array_equal(capi.cluster_poses(both, [1.0, 0.5], ang, 1.0), O.ref_cluster_poses(both, [1.0, 0.5], ang, 1.0))


def test_cluster_golden_fixture():
    g = np.load(GOLD)
    for k in range(int(g["n_cases"])):
        got = capi.cluster_poses(g[f"hyp{k}"], g[f"scores{k}"], float(g[f"angle{k}"]), float(g[f"dist{k}"]), g[f"sym{k}"])
        assert np.array_equal(got, g[f"keep{k}"])


def test_cluster_edge_cases():
    I = np.eye(4, dtype=np.float32)[None]
    assert list(capi.cluster_poses(I, [0.3], 30, 0.015)) == [0]
    same = np.repeat(I, 5, 0)
    assert list(capi.cluster_poses(same, [0.1, 0.5, 0.5, 0.2, 0.5], 30, 0.015)) == [1]       # best score, lowest id survives alone
    far = same.copy(); far[:, 0, 3] = np.arange(5) * 0.1
    assert list(capi.cluster_poses(far, [0.1, 0.5, 0.5, 0.2, 0.5], 30, 0.015)) == [1, 2, 4, 3, 0]
    # the reference's geodesic distance uses trace(R1 * R2), not R1^T R2: two equal 90-degree rotations are "far" (:29-32)
    Rz = np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    two = np.stack([Rz, Rz])
    assert list(capi.cluster_poses(two, [0.5, 0.4], 30, 0.015)) == [0]                        # merged by the Euler test
    assert list(capi.cluster_poses(two, [0.5, 0.4], 30, 0.015, (-1.0, -1.0, -1.0))) == [0]
