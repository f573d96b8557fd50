"""CPU tests of the physics-pruning oracle (oracle/hop_oracle_sdf.c): pinned against the reference's OWN libigl compiled from
where it lies (oracle/_ref/libhop_ref.so: igl::signed_distance, pseudonormal), on closed meshes with big and small faces
(both branches of igl::pseudonormal_test), and the reject decision on grasp scenes against the committed golden fixture.

Bars: |S| within 1e-7 m of igl (float, same Ericson region walk); the sign equal except where igl's AABB tree returned ANOTHER
face of an exact distance tie (a point nearest to a shared edge / vertex of small faces, where the reference takes whichever
face its tree met first); such points must be < 0.1 % and the two faces must really tie."""
import os

import numpy as np
import pytest

from hop_b200 import synth
from oracle import cpu_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "collision_golden.npz")


def _has_ref():
    try:
        return hasattr(O.ref(), "hop_ref_signed_distance")
    except Exception:
        return False


needs_ref = pytest.mark.skipif(not _has_ref(), reason="oracle/_ref/libhop_ref.so (compiled reference) not built")


@needs_ref
@pytest.mark.parametrize("name,level", [("ellipse", 1), ("ellipse", 3), ("cuboid", 1), ("cuboid", 3), ("cylinder", 2), ("tless", 2), ("tless", 3)])
def test_signed_distance_pinned_to_igl(name, level):
    rng = np.random.default_rng(7 + level)
    V, F = synth.make_mesh(name, level)
    R = synth.random_rotation(rng)
    Vt = (V @ R.T + rng.normal(0, 0.1, 3)).astype(np.float32)
    pts = rng.uniform(Vt.min(0) - 0.02, Vt.max(0) + 0.02, (8000, 3)).astype(np.float32)
    k = min(50, len(Vt))
    pts = np.concatenate([pts, Vt[:k], (Vt[F[:50, 0]] + Vt[F[:50, 1]]) / 2]).astype(np.float32)  # on vertices and edges too
    S, I, Cp = O.signed_distance(pts, Vt, F)
    Sr, Ir, Cr, _ = O.ref_signed_distance(pts, Vt, F)
    # a point lying exactly ON the mesh is "out of bounds" for igl (sqrd <= low_sqr_d = 0, signed_distance.cpp:158): NaN in both
    # (edge midpoints may round to a distance of 0 in one and 1e-9 in the other: there the non-NaN one must be ~0)
    assert np.all(np.isnan(S[8000:8000 + k])) and np.all(np.isnan(Sr[8000:8000 + k]))
    one = np.isnan(S) != np.isnan(Sr)
    assert np.all(np.abs(np.where(np.isnan(S), Sr, S)[one]) < 1e-7)
    ok = ~np.isnan(S) & ~np.isnan(Sr)
    assert np.abs(np.abs(S[ok]) - np.abs(Sr[ok])).max() < 1e-7
    off = ok & (np.abs(S) > 1e-6)                # next to the surface itself the sign is noise in both
    bad = off & (np.sign(S) != np.sign(Sr))
    assert bad.mean() < 1e-3
    assert np.all(I[bad] != Ir[bad])             # every sign difference comes from a face tie ...
    assert np.all(np.abs(np.linalg.norm(pts[bad] - Cp[bad], axis=1) - np.linalg.norm(pts[bad] - Cr[bad], axis=1)) < 1e-7)
    same = I == Ir
    assert np.array_equal(np.sign(S[same & off]), np.sign(Sr[same & off]))   # same face -> same pseudonormal rule -> same sign


def test_signed_distance_known_answers():
    """analytic cases on the unit-ish cuboid: face, edge and vertex regions, inside"""
    V, F = synth.make_mesh("cuboid", 1)           # 0.08 x 0.05 x 0.03, faces above MIN_DOUBLE_AREA
    pts = np.array([[0.06, 0, 0], [0, 0, 0], [0.05, 0.035, 0], [0.05, 0.035, 0.025], [0.039, 0, 0]], np.float32)
    S, _, _ = O.signed_distance(pts, V, F)
    want = [0.02, -0.015, np.hypot(0.01, 0.01), np.sqrt(3) * 0.01, -0.001]
    assert np.allclose(S, want, atol=1e-7)


def test_meshes_closed_and_outward():
    for name in ("ellipse", "cuboid", "cylinder", "tless"):
        V, F = synth.make_mesh(name, 2)
        e = np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), 1)
        _, cnt = np.unique(e, axis=0, return_counts=True)
        assert np.all(cnt == 2)
        vol = np.einsum("ij,ij->i", V[F[:, 0]].astype(float), np.cross(V[F[:, 1]], V[F[:, 2]]).astype(float)).sum() / 6
        assert vol > 0


@needs_ref
def test_reject_steps_pinned_to_igl_on_moved_mesh():
    """the decision's distances, recomputed the reference's way: transformVertices moves the mesh into the hypothesis' frame
    (float), igl::signed_distance queries the finger cloud there (SDFchecker.cpp:22-33,115-134)"""
    case = synth.make_collision_case("ellipse", H=24, seed=3)
    keep, reason, diag = O.reject_by_collision(case)
    cam2hb = case["params"]["cam2handbase"]
    checked = 0
    for h in range(len(keep)):
        M = (cam2hb @ case["poses"][h]).astype(np.float32)
        Vt = (case["obj_V"] @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
        for k in range(4):
            if diag[h, 2 + k] > 1e30:
                continue
            Sr = O.ref_signed_distance(case["finger_pts"][k], Vt, case["obj_F"])[0]
            assert abs(Sr.min() - diag[h, 2 + k]) < 2e-7
            checked += 1
    assert checked >= 20


def test_reject_by_collision_golden():
    """committed fixture (tests/golden/make_collision_golden.py): decisions and distances of the restatement"""
    g = np.load(GOLD)
    for name in ("ellipse", "cuboid", "tless"):
        case = synth.make_collision_case(name, H=int(g[f"{name}_H"]), seed=int(g[f"{name}_seed"]))
        keep, reason, diag = O.reject_by_collision(case)
        assert np.array_equal(reason, g[f"{name}_reason"])
        assert np.array_equal(keep, (g[f"{name}_reason"] == 0).astype(np.int32))
        fin = g[f"{name}_diag"] < 1e30
        assert np.array_equal(fin, diag < 1e30)
        assert np.abs(diag[fin] - g[f"{name}_diag"][fin]).max() < 1e-6
    # every branch of the decision is exercised by the fixture
    allr = np.concatenate([g[f"{n}_reason"] for n in ("ellipse", "cuboid", "tless")])
    assert set(np.unique(allr)) >= {0, 1, 3, 4, 5}


def test_reject_disabled_fingers_and_integer_ratio():
    case = synth.make_collision_case("ellipse", H=64, seed=5, disabled=(0,))     # finger_1_1 off: finger 1 is skipped entirely
    keep, reason, diag = O.reject_by_collision(case)
    assert np.all(diag[:, 2] > 1e30) and np.all(diag[:, 3] > 1e30)
    # num_inside / P.rows() is an integer division in the reference: reason 6 needs EVERY model point inside a finger mesh
    assert not np.any(reason == 6)
