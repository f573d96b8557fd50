"""GPU parity tests of hop_remove_hand_points (HandT42::removeSurroundingPointsAndAssignProbability, Hand.cpp:781-888) through
the C ABI against its restatement (oracle/hop_oracle_hand.c).  Bars: the kept set, the order, positions and normals BIT-EXACT
(same unfused float operations); the confidence 1 - exp(-lambda d) within 2e-7 (device expf vs libm expf)."""
import numpy as np
import pytest

from hop_b200 import synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _run(ctx, case, links=None, kinds=None):
    links = case["links"] if links is None else links
    kinds = case["kinds"] if kinds is None else kinds
    p = ctx.hand_removal_params(case["handbase_in_cam"], case["finger_1_2_in_handbase"], case["finger_2_2_in_handbase"], case["min_z"], case["near_hand_dist"])
    scene = ctx.upload_cloud(case["scene_xyz"], case["scene_nrm"])
    lc = [ctx.upload_cloud(l) if len(l) else None for l in links]
    out = ctx.remove_hand_points(scene, lc, kinds, p)
    got = out.download()
    want = O.remove_hand_points(case["scene_xyz"], case["scene_nrm"], links, kinds, p)
    for c in [scene, out] + [c for c in lc if c is not None]:
        c.free()
    return got, want


@pytest.mark.parametrize("seed,n", [(4, 3000), (5, 20000), (6, 777)])
def test_remove_hand_points_matches_oracle(ctx, seed, n):
    case = synth.make_hand_removal_case(seed=seed, n_scene=n)
    (x, nr, c), (ox, on, oc) = _run(ctx, case)
    assert len(x) == len(ox) and 0.2 * n < len(x) < 0.8 * n
    assert np.array_equal(x, ox) and np.array_equal(nr, on)
    assert np.abs(c - oc).max() < 2e-7
    assert (c >= 0.8).sum() > 0.1 * n                       # what setCurScene keeps: the object and the clutter


def test_remove_hand_points_edge_cases(ctx):
    case = synth.make_hand_removal_case(seed=8, n_scene=500)
    # no links at all: only the outer-side rule acts, min_dist stays 1.0
    (x, nr, c), (ox, on, oc) = _run(ctx, case, links=[], kinds=[])
    assert np.array_equal(x, ox) and np.allclose(c, 1 - np.exp(-231.04906018664843), atol=2e-7)
    # an empty link cloud in the middle is skipped like FLANN's empty search
    links = list(case["links"]); links[2] = links[2][:0]
    (x, nr, c), (ox, on, oc) = _run(ctx, case, links=links)
    assert np.array_equal(x, ox) and np.abs(c - oc).max() < 2e-7
    # empty scene
    empty = dict(case, scene_xyz=case["scene_xyz"][:0], scene_nrm=case["scene_nrm"][:0])
    (x, nr, c), _ = _run(ctx, empty)
    assert len(x) == 0
    # everything is hand: a scene made of the link points themselves
    hic = case["handbase_in_cam"]
    pts = np.concatenate(case["links"]) @ hic[:3, :3].T + hic[:3, 3]
    allhand = dict(case, scene_xyz=pts.astype(np.float32), scene_nrm=np.tile([0, 0, 1.0], (len(pts), 1)).astype(np.float32))
    (x, nr, c), (ox, _, _) = _run(ctx, allhand)
    assert len(x) == len(ox) == 0


def _height_case(seed, true_height):
    """the hand cloud (links of the removal fixture, outward normals) and a hand-region scene that shows the hand `true_height`
    higher along the hand base's z than the kinematics say"""
    rng = np.random.default_rng(seed)
    pts, nrm = [], []
    for size, off in (((0.08, 0.10, 0.03), (-0.06, 0.0, 0.0)), ((0.05, 0.012, 0.02), (-0.14, -0.045, 0.0)), ((0.05, 0.012, 0.02), (-0.14, 0.045, 0.0))):
        p, n = synth._cuboid(rng, 500, *size)
        pts.append(p + off); nrm.append(n)
    hand_xyz, hand_nrm = np.concatenate(pts).astype(np.float32), np.concatenate(nrm).astype(np.float32)
    seen = rng.choice(len(hand_xyz), 900, replace=False)
    scene = hand_xyz[seen] + [0, 0, true_height] + rng.normal(0, 0.0005, (900, 3))
    snrm = hand_nrm[seen] + rng.normal(0, 0.05, (900, 3))
    clutter = rng.uniform([-0.2, -0.1, -0.06], [0.0, 0.1, 0.06], (600, 3))
    cn = rng.normal(size=(600, 3)); cn /= np.linalg.norm(cn, axis=1, keepdims=True)
    return hand_xyz, hand_nrm, np.concatenate([scene, clutter]).astype(np.float32), np.concatenate([snrm, cn]).astype(np.float32)


@pytest.mark.parametrize("true_height", [0.01, -0.02, 0.0])
def test_adjust_hand_height_matches_oracle(ctx, true_height):
    hx, hn, sx, sn = _height_case(3, true_height)
    hand, scene = ctx.upload_cloud(hx, hn), ctx.upload_cloud(sx, sn)
    counts, best = ctx.adjust_hand_height(hand, scene)
    ocounts, obest = O.adjust_hand_height(hx, hn, sx, sn, ctx.TRIAL_HEIGHTS)
    assert np.array_equal(counts, ocounts) and best == obest
    assert abs(ctx.TRIAL_HEIGHTS[best] - true_height) < 1e-6 and counts[best] > 300
    hand.free(); scene.free()


def test_adjust_hand_height_nothing_matches(ctx):
    hx, hn, sx, sn = _height_case(4, 0.0)
    hand, scene = ctx.upload_cloud(hx, hn), ctx.upload_cloud((sx + [0, 0, 1.0]).astype(np.float32), sn)
    counts, best = ctx.adjust_hand_height(hand, scene)
    assert best == -1 and counts.sum() == 0            # the reference keeps _handbase_in_cam (best_height stays 0)
    hand.free(); scene.free()
