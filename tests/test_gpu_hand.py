"""GPU parity tests of K1 (hop_hand_overlap) against the objFuncPSO restatement, through the C ABI.

Bars: the integer part of the objective (matched finger points, branch taken) must agree exactly -- the cost of the
gap-penalty, no-match and match-only branches is then BIT-identical (same float/double operation order); the two branches
that contain the float sum over outer points are compared within 1e-5 relative (the reference sums sequentially, the
kernel in a fixed tree order)."""
import numpy as np
import pytest

from hop_b200 import hand, synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _run(ctx, case, thetas, use_lookup=True):
    prop = hand.FingerProperty(case["finger_xyz"], case["scalars"]["num_division"])
    p = hand.finger_params(prop, case["scalars"])
    finger = ctx.upload_cloud(case["finger_xyz"], case["finger_nrm"])
    scene = ctx.upload_cloud(case["scene_xyz"], case["scene_nrm"])
    lookup = ctx.upload_cloud(case["lookup_xyz"], case["lookup_nrm"]) if use_lookup else None
    nosw = ctx.upload_cloud(case["noswivel_xyz"], case["noswivel_nrm"])
    got, best = ctx.hand_overlap(finger, scene, nosw, p, thetas, lookup)
    ref, detail = O.hand_overlap(p, case["finger_xyz"], case["finger_nrm"], case["scene_xyz"],
                                 case["lookup_nrm"] if use_lookup else case["scene_nrm"], case["noswivel_xyz"], thetas, with_detail=True)
    for c in (finger, scene, nosw, lookup):
        if c is not None:
            c.free()
    return got, best, ref, detail


def _check(got, best, ref, detail):
    branch = detail[:, 3].astype(int)
    exact = np.isin(branch, [0, 1, 4])
    # branches without the outer float sum in the cost: bit-identical
    assert np.array_equal(got[exact], ref[exact]), np.nonzero(got[exact] != ref[exact])[0][:10]
    rest = ~exact
    # the penalty w * exp(1000 * avg) (branch 3) amplifies the rounding of avg by 1000 * avg: scale the bar by the penalty
    with np.errstate(invalid="ignore", divide="ignore"):
        avg = np.where(detail[:, 1] > 0, detail[:, 2] / np.maximum(detail[:, 1], 1), 0.0)
    pen = np.where(branch == 3, np.exp(1000 * avg) * (1 + 1000 * avg), 0.0)
    assert np.all(np.abs(got[rest] - ref[rest]) <= 1e-5 * (np.maximum(np.abs(ref[rest]), 1.0) + pen[rest]))
    assert best == int(np.argmin(ref)) or got[best] == ref.min()


@pytest.mark.parametrize("variant", list(synth.hand_variants()))
def test_hand_overlap_matches_oracle(ctx, variant):
    case = synth.hand_problem(variant, seed=5)
    thetas = np.deg2rad(np.linspace(0, 120, 1537))
    got, best, ref, detail = _run(ctx, case, thetas)
    _check(got, best, ref, detail)
    assert len(set(detail[:, 3].astype(int))) >= 2


def test_hand_overlap_without_lookup_cloud_and_other_seeds(ctx):
    for seed in (11, 12):
        case = synth.hand_problem("left", seed=seed)
        thetas = np.deg2rad(np.linspace(0, 90, 1000))
        got, best, ref, detail = _run(ctx, case, thetas, use_lookup=False)
        _check(got, best, ref, detail)
        assert abs(np.rad2deg(thetas[best]) - np.rad2deg(case["theta_true"])) < 2.0


def test_hand_overlap_streams_a_large_scene(ctx):
    """a no-swivel scene of several shared-memory stages (> 2048 points each) and a ragged tail"""
    case = synth.make_hand_case(seed=21, n_hand=9001, n_finger=777)
    thetas = np.deg2rad(np.linspace(0, 60, 257))
    got, best, ref, detail = _run(ctx, case, thetas)
    assert len(case["noswivel_xyz"]) > 3 * 2048
    _check(got, best, ref, detail)


def test_hand_overlap_edge_cases(ctx):
    case = synth.hand_problem("left", seed=3)
    got, best, ref, detail = _run(ctx, case, np.zeros(0))
    assert got.shape == (0,) and best == -1
    got, best, ref, detail = _run(ctx, case, np.array([case["theta_true"]]))       # a single state
    _check(got, best, ref, detail)
    # ties: identical states -> the lowest index wins the arg-min
    got, best, ref, detail = _run(ctx, case, np.full(70, case["theta_true"]))
    assert best == 0 and np.all(got == got[0])
    # empty no-swivel scene: avg = 0/0 = NaN -> no outer penalty (Hand.cpp:153)
    empty = dict(case)
    empty["noswivel_xyz"], empty["noswivel_nrm"] = np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)
    got, best, ref, detail = _run(ctx, empty, np.deg2rad(np.linspace(0, 40, 64)))
    assert np.array_equal(got, ref)


def test_hand_matcher_mirror(ctx):
    """Hand::matchOneComponentPSO through the host mirror: success, angle close to the truth, failure semantics."""
    case = synth.hand_problem("left", seed=31)
    hm = hand.HandMatcher(ctx, n_states=4096)
    hm.addComponent("finger_1_2", case["finger_xyz"], case["finger_nrm"])
    hm.setCurScene(case["scene_xyz"], case["scene_nrm"], case["noswivel_xyz"], case["noswivel_nrm"], case["lookup_xyz"], case["lookup_nrm"])
    assert hm.matchOneComponentPSO("finger_1_2", 0, 90, case["scalars"], least_match=5)
    assert abs(np.rad2deg(hm.angle) - np.rad2deg(case["theta_true"])) < 1.5 and hm._component_status["finger_1_2"]
    assert not hm.matchOneComponentPSO("finger_1_2", 0, 90, case["scalars"], least_match=1e6)   # PSO "failed": finger disabled
    assert np.array_equal(hm._tf_self["finger_1_2"], np.eye(4, dtype=np.float32)) and not hm._component_status["finger_1_2"]
