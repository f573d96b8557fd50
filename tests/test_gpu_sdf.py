"""GPU parity tests of the physics pruning step (hop_sdf_query, hop_reject_by_collision) through the C ABI, against the
restatement of igl::signed_distance / PoseEstimator::rejectByCollisionOrNonTouching (oracle/hop_oracle_sdf.c, pinned to the
reference's compiled libigl by tests/test_sdf_oracle.py) and, where the prebuilt oracle/_ref travels, against that libigl itself.

Bars (float): |S| within 2e-7 m; the sign equal wherever the closest face is the same (same pseudonormal rule) and < 0.1 %
different overall (face ties); decisions identical except for hypotheses with a tested distance within 1e-6 m of its threshold."""
import numpy as np
import pytest

from hop_b200 import synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _upload_case(ctx, case):
    obj = ctx.upload_mesh(case["obj_V"], case["obj_F"])
    fm = [ctx.upload_mesh(v, f) for v, f in zip(case["finger_V"], case["finger_F"])]
    status = case["params"]["finger_status"]
    fc = [ctx.upload_cloud(p) if status[k] else None for k, p in enumerate(case["finger_pts"])]
    scene, hand, model = ctx.upload_cloud(case["scene_xyz"]), ctx.upload_cloud(case["hand_xyz"]), ctx.upload_cloud(case["model_xyz"])
    return obj, fm, fc, scene, hand, model


def _oracle_case(case):
    c = dict(case)
    status = case["params"]["finger_status"]
    c["finger_pts"] = [p if status[k] else p[:0] for k, p in enumerate(case["finger_pts"])]   # disabled links are not in the map
    return c


@pytest.mark.parametrize("name,level", [("ellipse", 1), ("ellipse", 3), ("cuboid", 1), ("cuboid", 3), ("cylinder", 2), ("tless", 3)])
def test_sdf_query_matches_oracle(ctx, name, level):
    rng = np.random.default_rng(11 + level)
    V, F = synth.make_mesh(name, level)
    mesh = ctx.upload_mesh(V, F)
    pts = rng.uniform(V.min(0) - 0.02, V.max(0) + 0.02, (6000, 3)).astype(np.float32)
    pts = np.concatenate([pts, V[: min(40, len(V))]]).astype(np.float32)
    got = ctx.sdf_query(mesh, pts, want_faces=True)
    S, I, _ = O.signed_distance(pts, V, F)
    g = got["S"][0]
    assert np.array_equal(np.isnan(g[6000:]), np.isnan(S[6000:])) and np.all(np.isnan(g[6000:]))   # exactly on the mesh: NaN like igl
    ok = ~np.isnan(g) & ~np.isnan(S)
    assert ok.sum() >= 6000
    assert np.abs(np.abs(g[ok]) - np.abs(S[ok])).max() < 2e-7
    same = ok & (got["I"][0] == I) & (np.abs(S) > 1e-6)
    assert same.mean() > 0.98
    assert np.array_equal(np.sign(g[same]), np.sign(S[same]))
    assert (np.sign(g[ok]) != np.sign(S[ok])).mean() < 1e-3
    assert abs(got["min"][0] - np.nanmin(S)) < 2e-7 and abs(got["max"][0] - np.nanmax(S)) < 2e-7
    assert abs(int(got["n_inside"][0]) - int((S < 0).sum())) <= 6
    mesh.free()


def test_sdf_query_placements_equal_moved_mesh(ctx):
    """the inverse-placement trick: querying T^-1 p against the resting mesh == querying p against the mesh moved by T
    (what SDFchecker::transformMesh + igl do per hypothesis)"""
    rng = np.random.default_rng(5)
    V, F = synth.make_mesh("tless", 2)
    mesh = ctx.upload_mesh(V, F)
    pts = rng.uniform(-0.08, 0.08, (500, 3)).astype(np.float32)
    T = np.stack([np.eye(4) for _ in range(5)])
    for h in range(5):
        T[h, :3, :3] = synth.random_rotation(rng)
        T[h, :3, 3] = rng.normal(0, 0.02, 3)
    got = ctx.sdf_query(mesh, pts, point_transforms=np.linalg.inv(T))
    for h in range(5):
        Vt = (V @ T[h, :3, :3].T + T[h, :3, 3]).astype(np.float32)
        S, _, _ = O.signed_distance(pts, Vt, F)
        assert np.abs(got["S"][h] - S).max() < 5e-7
        assert abs(got["min"][h] - S.min()) < 5e-7 and int(got["n_inside"][h]) == int((S < 0).sum())
    mesh.free()


def test_sdf_query_edge_cases(ctx):
    V, F = synth.make_mesh("cuboid", 1)
    mesh = ctx.upload_mesh(V, F)
    r = ctx.sdf_query(mesh, np.zeros((0, 3), np.float32))
    assert r["min"][0] > 1e30 and r["max"][0] < -1e30 and r["n_inside"][0] == 0
    r = ctx.sdf_query(mesh, np.array([[0, 0, 0], [0.06, 0, 0]], np.float32))
    assert np.allclose(r["S"][0], [-0.015, 0.02], atol=1e-7) and r["n_inside"][0] == 1
    mesh.free()
    with pytest.raises(Exception):
        ctx.upload_mesh(V, F + 100)            # face index out of range


def _compare_decisions(case, got, want):
    keep, reason, diag = got
    okeep, oreason, odiag = want
    p = case["params"]
    thr = [p["inside_ob_dist"], p["collision_dist"]] + [p["collision_dist"], p["non_touch_dist"]] * 0
    # distances the oracle evaluated must agree (the kernel also evaluates the later fingers of a rejected hypothesis)
    fin = odiag < 1e30
    assert np.all(diag[fin] < 1e30)
    assert np.abs(diag[fin] - odiag[fin]).max() < 5e-7
    diff = np.nonzero(reason != oreason)[0]
    for h in diff:                                # only threshold ties may differ
        d = odiag[h][odiag[h] < 1e30]
        t = np.array([p["inside_ob_dist"], p["collision_dist"], p["non_touch_dist"], p["collision_finger_dist"]])
        assert np.abs(d[:, None] - t[None, :]).min() < 1e-6, (h, reason[h], oreason[h])
    assert len(diff) <= max(1, len(reason) // 200)
    assert np.array_equal(keep, (reason == 0).astype(np.int32))
    del thr


@pytest.mark.parametrize("name,seed", [("ellipse", 21), ("cuboid", 22), ("tless", 23), ("cylinder", 24)])
def test_reject_by_collision_matches_oracle(ctx, name, seed):
    case = synth.make_collision_case(name, H=192, seed=seed)
    obj, fm, fc, scene, hand, model = _upload_case(ctx, case)
    got = ctx.reject_by_collision(obj, fm, fc, scene, hand, model, case["poses"], case["params"])
    _compare_decisions(case, got, O.reject_by_collision(_oracle_case(case)))


def test_reject_by_collision_golden(ctx):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "collision_golden.npz"))
    for name in ("ellipse", "cuboid", "tless"):
        case = synth.make_collision_case(name, H=int(g[f"{name}_H"]), seed=int(g[f"{name}_seed"]))
        obj, fm, fc, scene, hand, model = _upload_case(ctx, case)
        got = ctx.reject_by_collision(obj, fm, fc, scene, hand, model, case["poses"], case["params"])
        _compare_decisions(case, got, ((g[f"{name}_reason"] == 0).astype(np.int32), g[f"{name}_reason"], g[f"{name}_diag"]))


def test_reject_disabled_links_and_missing_inputs(ctx):
    case = synth.make_collision_case("ellipse", H=96, seed=31, disabled=(0, 3))
    obj, fm, fc, scene, hand, model = _upload_case(ctx, case)
    got = ctx.reject_by_collision(obj, fm, fc, scene, hand, model, case["poses"], case["params"])
    _compare_decisions(case, got, O.reject_by_collision(_oracle_case(case)))
    # no scene cloud, no hand cloud, no finger meshes: only the finger-cloud steps remain
    got = ctx.reject_by_collision(obj, None, fc, None, None, model, case["poses"], case["params"])
    c2 = _oracle_case(case)
    c2["scene_xyz"], c2["hand_xyz"] = case["scene_xyz"][:0], case["hand_xyz"][:0]
    c2["finger_V"], c2["finger_F"] = [v[:0] for v in case["finger_V"]], [f[:0] for f in case["finger_F"]]
    _compare_decisions(case, got, O.reject_by_collision(c2))
    # H = 0
    k, r, d = ctx.reject_by_collision(obj, fm, fc, scene, hand, model, case["poses"][:0], case["params"])
    assert len(k) == 0


def test_pose_estimator_reject_method(ctx):
    """the host mirror: PoseEstimator.rejectByCollisionOrNonTouching keeps the survivors in order"""
    from hop_b200.pose_estimator import PoseEstimator
    case = synth.make_collision_case("cuboid", H=64, seed=41)
    pe = PoseEstimator(ctx, {})
    mx, mn = synth.make_model("cuboid", 500, seed=41 + 11)
    pe.setModel(mx, mn)
    pe.registerMesh(case["obj_V"], case["obj_F"], "object")
    for k, n in enumerate(pe.FINGERS):
        pe.registerMesh(case["finger_V"][k], case["finger_F"][k], n)
    pe.setPoseHypos(case["poses"])
    hand = dict(component_status={n: True for n in pe.FINGERS}, finger_clouds={n: case["finger_pts"][k] for k, n in enumerate(pe.FINGERS)},
                hand_cloud=case["hand_xyz"], handbase_in_cam=np.linalg.inv(case["params"]["cam2handbase"]), cloud_withouthand=case["scene_xyz"])
    pe.rejectByCollisionOrNonTouching(hand, {"collision_thres": 0.4})
    ids = [h._id for h in pe._pose_hypos]
    assert ids == sorted(ids) and 0 < len(ids) < 64
    assert np.array_equal(np.nonzero(pe._reject_reason == 0)[0], ids)


def test_reject_full_size_properties(ctx):
    """BASELINE C2-sized batch (1024 hypotheses, 10 k-point model): decisions are invariant under a permutation of the
    hypotheses and of the points of every cloud (the reductions are min / count), and the true pose survives"""
    case = synth.make_collision_case("ellipse", H=1024, seed=51, n_model=10000, mesh_level=3)
    case["poses"][0] = case["gt"]
    obj, fm, fc, scene, hand, model = _upload_case(ctx, case)
    k1, r1, d1 = ctx.reject_by_collision(obj, fm, fc, scene, hand, model, case["poses"], case["params"])
    assert k1[0] == 1
    rng = np.random.default_rng(1)
    perm = rng.permutation(1024)
    c2 = dict(case)
    c2["finger_pts"] = [p[rng.permutation(len(p))] for p in case["finger_pts"]]
    c2["model_xyz"] = case["model_xyz"][rng.permutation(len(case["model_xyz"]))]
    obj2, fm2, fc2, scene2, hand2, model2 = _upload_case(ctx, c2)
    k2, r2, d2 = ctx.reject_by_collision(obj2, fm2, fc2, scene2, hand2, model2, case["poses"][perm], case["params"])
    assert np.array_equal(r2, r1[perm]) and np.array_equal(d2, d1[perm])
