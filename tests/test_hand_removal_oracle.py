"""CPU tests of the hand-point removal restatement (oracle/hop_oracle_hand.c: hop_oracle_remove_hand_points, following
HandT42::removeSurroundingPointsAndAssignProbability, Hand.cpp:781-888).  PARITY UNPINNED against PCL/FLANN (absent); the
rules are checked on hand-made points and against a numpy restatement with scipy's exact nearest neighbour."""
import ctypes as C

import numpy as np
from scipy.spatial import cKDTree

from hop_b200 import synth
from oracle import cpu_oracle as O


def _params(case):
    p = O.HandRemovalParams()
    hic = case["handbase_in_cam"]
    for name, M in (("cam_in_handbase", np.linalg.inv(hic)), ("handbase_in_cam", hic), ("handbase_in_finger_1_2", np.linalg.inv(case["finger_1_2_in_handbase"])),
                    ("handbase_in_finger_2_2", np.linalg.inv(case["finger_2_2_in_handbase"]))):
        getattr(p, name)[:] = np.asarray(M, np.float32).T.reshape(-1).tolist()
    p.min_z = case["min_z"]
    p.dist_thres_sq = float(np.float32(case["near_hand_dist"]) ** 2)
    return p


def test_rules_on_hand_made_points():
    eye = np.eye(4, dtype=np.float32)
    far = np.eye(4, dtype=np.float32); far[:3, 3] = [0, -10, 0]         # distal frames far away on the -y side: y > 0 for every point, the outer-side rule never fires
    case = dict(handbase_in_cam=eye, finger_1_2_in_handbase=far, finger_2_2_in_handbase=far, min_z=0.0, near_hand_dist=0.003)
    link = np.array([[0, 0, 0], [0.1, 0, 0]], np.float32)
    pts = np.array([[0, 0, 0.002],        # within 3 mm of a link point: hand
                    [0.001, 0, 0.0045],   # 4.6 mm away, but planar 1 mm and |dz| <= 5 mm: hand
                    [0.001, 0, 0.0055],   # |dz| > 5 mm: kept
                    [0.05, 0.02, 0],      # far: kept
                    [0.0, 0.0039, 0]], np.float32)   # 3.9 mm: kept for kind 0, hand for kind 1 (5 mm)
    nrm = np.tile([0, 0, 1.0], (5, 1)).astype(np.float32)
    x, n, c = O.remove_hand_points(pts, nrm, [link], [0], _params(case))
    assert np.array_equal(x, pts[[2, 3, 4]]) and np.array_equal(n, nrm[:3])
    d = np.array([np.float32(np.sqrt(np.float32(0.001) ** 2 + np.float32(0.0055) ** 2)), np.sqrt(np.float32(0.05) ** 2 + np.float32(0.02) ** 2), 0.0039])
    assert np.allclose(c, 1 - np.exp(-231.04906018664843 * d), atol=2e-6)
    x1, _, _ = O.remove_hand_points(pts, nrm, [link], [1], _params(case))
    assert np.array_equal(x1, pts[[2, 3]])
    x2, _, c2 = O.remove_hand_points(pts, nrm, [link], [2], _params(case))      # 20 mm: only the far point survives
    assert np.array_equal(x2, pts[[3]])
    # no links: everything is kept with min_dist = 1.0
    x3, _, c3 = O.remove_hand_points(pts, nrm, [], [], _params(case))
    assert len(x3) == 5 and np.allclose(c3, 1 - np.exp(-231.04906018664843))
    # outer side of a distal link: y < 0 and z >= min_z in the link frame
    near = np.eye(4, dtype=np.float32); near[:3, 3] = [0.05, 0.03, -0.01]
    case2 = dict(case, finger_1_2_in_handbase=near)
    x4, _, _ = O.remove_hand_points(pts, nrm, [link], [0], _params(case2))
    keep = [i for i in (2, 3, 4) if not ((pts[i, 1] - 0.03 < 0) and (pts[i, 2] + 0.01 >= 0.0))]
    assert np.array_equal(x4, pts[keep])


def test_against_numpy_restatement_with_exact_nn():
    case = synth.make_hand_removal_case(seed=4, n_scene=3000)
    x, n, c = O.remove_hand_points(case["scene_xyz"], case["scene_nrm"], case["links"], case["kinds"], _params(case))
    cih = np.linalg.inv(case["handbase_in_cam"]).astype(np.float32)
    hb = (case["scene_xyz"] @ cih[:3, :3].T + cih[:3, 3]).astype(np.float32)
    thr = {0: np.float32(0.003) ** 2, 1: np.float32(0.005 * 0.005), 2: np.float32(0.02 * 0.02)}
    trees = [cKDTree(l.astype(np.float64)) for l in case["links"]]
    keep, conf = [], []
    for i, p in enumerate(hb):
        near, md = False, 1.0
        for k, t in enumerate(trees):
            d, j = t.query(p.astype(np.float64))
            md = min(md, d)
            q = case["links"][k][j]
            if d * d <= thr[case["kinds"][k]] or ((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 <= thr[case["kinds"][k]] and abs(p[2] - q[2]) <= 0.005):
                near = True
                break
        if near:
            continue
        out = False
        for F in (case["finger_1_2_in_handbase"], case["finger_2_2_in_handbase"]):
            Fi = np.linalg.inv(F)
            q = Fi[:3, :3] @ p + Fi[:3, 3]
            out = out or (q[1] < 0 and q[2] >= case["min_z"])
        if not out:
            keep.append(i); conf.append(1 - np.exp(-231.04906018664843 * md))
    # the output went camera -> hand base -> camera: match it back to the input by nearest neighbour (1e-6 m)
    d_back, idx = cKDTree(case["scene_xyz"].astype(np.float64)).query(x.astype(np.float64))
    assert d_back.max() < 1e-6 and np.all(np.diff(idx) > 0)             # input order kept
    assert len(set(idx) ^ set(keep)) <= 2 and 0.2 * len(hb) < len(x) < 0.8 * len(hb)   # (a threshold can sit at float rounding)
    common = sorted(set(idx) & set(keep))
    cg = c[np.searchsorted(idx, common)]
    cw = np.array(conf)[np.searchsorted(np.array(keep), common)]
    assert np.abs(cg - cw).max() < 1e-5
    assert np.all((c >= 0) & (c <= 1)) and (c > 0.8).mean() > 0.2        # setCurScene keeps confidence >= 0.8: the object survives
