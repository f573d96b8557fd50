"""CPU checks of the K1 oracle (objFuncPSO restatement) and of the host mirror's FingerProperty."""
import ctypes as C

import numpy as np
import pytest

import hop_b200
from hop_b200 import hand, synth
from oracle import cpu_oracle as O


def _problem(seed, variant="left"):
    case = synth.hand_problem(variant, seed)
    prop = hand.FingerProperty(case["finger_xyz"], case["scalars"]["num_division"])
    return case, prop, hand.finger_params(prop, case["scalars"])


def test_finger_property_mirror_matches_restatement():
    for seed in range(3):
        case, prop, p = _problem(seed)
        q = hop_b200.FingerParams()
        bbox = O.finger_property(case["finger_xyz"], 10, q)
        assert np.array_equal(bbox, np.array([prop._min_x, prop._min_y, prop._min_z, prop._max_x, prop._max_y, prop._max_z], np.float32))
        assert (q.num_division, q.min_z, q.stride_z) == (p.num_division, p.min_z, p.stride_z)
        assert list(q.hist_min_y)[:10] == list(p.hist_min_y)[:10]
    # a sparse cloud leaves bins untouched: they take the next touched bin's column (Hand.cpp:213-235)
    xyz = np.array([[0, -1, 0.0], [0, -2, 0.35], [0, -3, 1.0]], np.float32)
    prop = hand.FingerProperty(xyz, 10)
    q = hop_b200.FingerParams()
    O.finger_property(xyz, 10, q)
    assert list(q.hist_min_y)[:10] == list(prop._hist_alongz[1]) == [-1, -2, -2, -2, -3, -3, -3, -3, -3, -3]


def _grid(case, p, thetas):
    return O.hand_overlap(p, case["finger_xyz"], case["finger_nrm"], case["scene_xyz"], case["lookup_nrm"], case["noswivel_xyz"], thetas,
                          with_detail=True)


@pytest.mark.parametrize("variant", ["left", "right", "nonormal"])
def test_objective_finds_the_true_angle(variant):
    case, prop, p = _problem(5, variant)
    thetas = np.deg2rad(np.linspace(0, 90, 721))
    cost, detail = _grid(case, p, thetas)
    best = int(np.argmin(cost))
    assert abs(np.rad2deg(thetas[best]) - np.rad2deg(case["theta_true"])) < 2.0
    assert -cost[best] > 200                                    # most of the 300 finger points match at the true angle
    gap = detail[:, 3] == 0
    assert gap.any() and np.all(cost[gap] >= 1e3)
    g = np.nonzero(gap)[0]
    first = cost[g[: len(g) // 3]]
    assert np.all(np.diff(first) >= 0)                          # "only use dist1 to make objective monotone"


def test_every_branch_is_reached():
    seen = set()
    thetas = np.deg2rad(np.linspace(0, 90, 361))
    for variant in synth.hand_variants():
        case, prop, p = _problem(5, variant)
        cost, detail = _grid(case, p, thetas)
        seen |= set(detail[:, 3].astype(int))
        if variant == "nogap":
            nomatch = detail[:, 3] == 1
            assert nomatch.any() and np.allclose(cost[nomatch], 100 - thetas[nomatch], atol=1e-5)
        if variant == "exp":
            e = detail[:, 3] == 3
            avg = (detail[e, 2] / detail[e, 1]).astype(np.float32)
            nm = -cost[e] + np.exp(avg * np.float32(1000))
            assert e.any() and np.all(nm > 0)
    assert seen == {0, 1, 2, 3, 4}


def test_objective_types_follow_the_reference():
    """num_match accumulates (1 + X) in double but is rounded to float after every match (Hand.cpp:101-121)."""
    case, prop, p = _problem(7)
    th = np.array([case["theta_true"]])
    cost, detail = O.hand_overlap(p, case["finger_xyz"], case["finger_nrm"], case["scene_xyz"], case["lookup_nrm"], case["noswivel_xyz"],
                                  th, with_detail=True)
    k = int(detail[0, 0])
    assert k > 20
    nm = np.float32(0)
    for _ in range(k):
        nm = np.float32(np.float64(nm) + (1 + th[0]))
    if int(detail[0, 3]) == 4:
        assert cost[0] == -np.float64(nm)
    # the neighbour's normal comes from the LOOKUP cloud at the kd-tree's index (Hand.cpp:94): other normals, other cost
    other = np.roll(case["lookup_nrm"], 7, axis=0)
    cost2 = O.hand_overlap(p, case["finger_xyz"], case["finger_nrm"], case["scene_xyz"], other, case["noswivel_xyz"], th)
    assert cost2[0] != cost[0]


def test_dense_grid_is_never_worse_than_the_swarm():
    """The reference's swarm (pso.hpp:146-351 schedule, 16 particles x (1 + 3) evaluations) can only visit points of
    [lb, ub]; the dense grid's optimum is at least as good up to the grid pitch."""
    case, prop, p = _problem(9)
    lb, ub = 0.0, np.deg2rad(90)
    args = (p, case["finger_xyz"], case["finger_nrm"], case["scene_xyz"], case["lookup_nrm"], case["noswivel_xyz"])
    grid = np.linspace(lb, ub, 4096)
    gcost = O.hand_overlap(*args, grid)
    L = O.lib()
    FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    f = FN(lambda x, _: float(O.hand_overlap(*args, np.array([x]))[0]))
    L.hop_oracle_pso_1d.restype = C.c_double
    L.hop_oracle_pso_1d.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_void_p, FN, C.c_void_p, C.POINTER(C.c_double)]
    worse = 0
    for seed in range(8):
        rng = np.random.default_rng(seed)
        ri, rc, rs = rng.random(15), rng.random(3 * 16), rng.random(3 * 16)
        bx = C.c_double()
        sw = L.hop_oracle_pso_1d(lb, ub, 15, 3, 0.1, 0.9, 0.0, ri.ctypes.data, rc.ctypes.data, rs.ctypes.data, f, None, C.byref(bx))
        assert lb <= bx.value <= ub
        worse += gcost.min() > sw + 1.5   # one match (1 + theta) of slack for the grid pitch
    assert worse == 0
