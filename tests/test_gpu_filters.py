"""GPU parity tests of the cloud filters of Hand::setCurScene (Hand.cpp:279-334) through the C ABI: hop_cloud_voxel_grid,
hop_cloud_transform, hop_cloud_pass_through, hop_cloud_radius_outlier_removal, hop_cloud_statistical_outlier_removal, against
numpy restatements of the PCL filters (oracle/cpu_oracle.py; float32 distances in FLANN's operation order, double statistics in
the reference's summation order).  PARITY UNPINNED against PCL itself (absent).  Bars: kept sets, order, positions and normals
BIT-EXACT; the voxel grid against the host C++ restatement's rule (leaf centroids in leaf-index order)."""
import numpy as np
import pytest

from hop_b200 import hand, synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _cloud(seed, n):
    case = synth.make_hand_removal_case(seed=seed, n_scene=n)
    return case, case["scene_xyz"], case["scene_nrm"]


@pytest.mark.parametrize("radius,min_n", [(0.02, 30), (0.04, 100), (0.005, 2)])
def test_radius_outlier_removal(ctx, radius, min_n):
    _, xyz, nrm = _cloud(3, 2500)
    c = ctx.upload_cloud(xyz, nrm)
    out = c.radius_outlier_removal(radius, min_n)
    x, n, w = out.download()
    keep = O.radius_outlier_removal(xyz, radius, min_n)
    assert 0 < keep.sum() < len(xyz)
    assert np.array_equal(x, xyz[keep]) and np.array_equal(n, nrm[keep])
    c.free(); out.free()


@pytest.mark.parametrize("k,mul", [(20, 2.0), (8, 1.0)])
def test_statistical_outlier_removal(ctx, k, mul):
    _, xyz, nrm = _cloud(4, 2000)
    c = ctx.upload_cloud(xyz, nrm)
    out = c.statistical_outlier_removal(k, mul)
    x, n, w = out.download()
    keep, dist = O.statistical_outlier_removal(xyz, k, mul)
    assert 0.5 * len(xyz) < keep.sum() < len(xyz)
    assert np.array_equal(x, xyz[keep]) and np.array_equal(n, nrm[keep])
    c.free(); out.free()


def test_transform_pass_through_voxel_grid(ctx):
    case, xyz, nrm = _cloud(5, 3000)
    conf = np.random.default_rng(0).random(len(xyz)).astype(np.float32)
    c = ctx.upload_cloud(xyz, nrm, conf)
    T = np.linalg.inv(case["handbase_in_cam"]).astype(np.float32)
    hb = c.transform(T)
    x, n, w = hb.download()
    ox, on = O.transform_cloud(T, xyz, nrm)
    assert np.array_equal(x, ox) and np.array_equal(n, on) and np.array_equal(w, conf)
    ps = hb.pass_through("x", -0.25, -0.1)
    px, pn, pw = ps.download()
    m = (ox[:, 0] >= np.float32(-0.25)) & (ox[:, 0] <= np.float32(-0.1))
    assert np.array_equal(px, ox[m]) and np.array_equal(pw, conf[m]) and 0 < m.sum() < len(m)
    vg = c.voxel_grid(0.003)
    vx, vn, vw = vg.download()
    # pcl::VoxelGrid: one centroid per occupied leaf, ordered by leaf index (x fastest)
    inv = np.float32(1.0) / np.float32(0.003)
    ijk = np.floor(xyz * inv).astype(np.int64)
    ijk -= ijk.min(0)
    div = ijk.max(0) + 1
    lin = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(lin, kind="stable")
    leaves, start = np.unique(lin[order], return_index=True)
    assert len(vx) == len(leaves)
    k = len(leaves) // 2
    members = order[start[k]:(start[k + 1] if k + 1 < len(start) else len(order))]
    s = np.zeros(3, np.float32)
    for i in members:
        s = (s + xyz[i]).astype(np.float32)
    assert np.array_equal(vx[k], (s / np.float32(len(members))).astype(np.float32))
    for cl in (c, hb, ps, vg):
        cl.free()


def test_filter_edge_cases(ctx):
    _, xyz, nrm = _cloud(6, 300)
    c = ctx.upload_cloud(xyz, nrm)
    e = c.pass_through("z", 100.0, 200.0)                      # nothing survives
    assert e.n == 0
    for f in (lambda cl: cl.radius_outlier_removal(0.01, 3), lambda cl: cl.statistical_outlier_removal(5, 1.0), lambda cl: cl.voxel_grid(0.01),
              lambda cl: cl.transform(np.eye(4)), lambda cl: cl.pass_through("x", -1, 1)):
        o = f(e)                                               # filters of an empty cloud are empty clouds
        assert o.n == 0
        o.free()
    few = ctx.upload_cloud(xyz[:4], nrm[:4])                   # fewer points than mean_k + 1: the neighbours there are
    o = few.statistical_outlier_removal(20, 2.0)
    assert 0 < o.n <= 4
    with pytest.raises(Exception):
        c.statistical_outlier_removal(100, 2.0)                # mean_k > 64
    for cl in (c, e, few, o):
        cl.free()


def test_set_cur_scene_chain(ctx):
    """Hand::setCurScene's whole chain on the device against the same chain of restatements"""
    case, xyz, nrm = _cloud(7, 12000)
    c = ctx.upload_cloud(xyz, nrm)
    got = hand.set_cur_scene(c, case["handbase_in_cam"])
    ds = c.voxel_grid(0.003)
    dx, dn, _ = ds.download()
    T = np.linalg.inv(case["handbase_in_cam"]).astype(np.float32)
    hx, hn = O.transform_cloud(T, dx, dn)
    k1 = O.radius_outlier_removal(hx, 0.02, 30)
    x1, n1 = hx[k1], hn[k1]
    k2 = O.radius_outlier_removal(x1, 0.04, 100)
    x2, n2 = x1[k2], n1[k2]
    k3, _ = O.statistical_outlier_removal(x2, 20, 2.0)
    x3, n3 = x2[k3], n2[k3]
    m = (x3[:, 0] >= np.float32(-0.25)) & (x3[:, 0] <= np.float32(-0.1))
    gx, gn, _ = got["scene_hand_region"].download()
    assert np.array_equal(gx, hx) and np.array_equal(gn, hn)
    gx, gn, _ = got["scene_hand_region_removed_noise"].download()
    assert np.array_equal(gx, x3) and np.array_equal(gn, n3) and 100 < len(x3) < len(hx)
    gx, gn, _ = got["scene_remove_swivel"].download()
    assert np.array_equal(gx, x3[m]) and len(gx) > 0
    for cl in list(got.values()) + [c, ds]:
        cl.free()


def test_handbase_icp_chain(ctx):
    """Hand::handbaseICP (Hand.cpp:677-763) through the mirror: the region cloud bit-exact against the restated filters, the ICP
    correction within the north-star tolerance of the oracle's runICP, the corrected hand base near the truth"""
    rng = np.random.default_rng(11)
    bp, bn = synth._cuboid(rng, 1500, 0.08, 0.10, 0.03)
    base_xyz, base_nrm = (bp + [-0.02, 0.0, -0.08]).astype(np.float32), bn.astype(np.float32)     # base_link cloud, hand-base frame
    # the kinematics put the hand base at hic_hat; the real one is off by a few millimetres and degrees
    hic_hat = np.eye(4); hic_hat[:3, :3] = synth.random_rotation(rng); hic_hat[:3, 3] = [0.05, -0.02, 0.5]
    true_off = np.eye(4); true_off[:3, :3] = synth._rot_from_rotvec(np.deg2rad([2.0, -3.0, 1.5])); true_off[:3, 3] = [0.004, -0.006, 0.005]
    hic_true = hic_hat @ np.linalg.inv(true_off)
    vis = np.nonzero(bn @ np.array([0.3, 0.2, 1.0]) > 0)[0]                   # the faces a camera above the hand sees
    sp = base_xyz[vis] + rng.normal(0, 0.0004, (len(vis), 3))
    clutter = rng.uniform([-0.3, -0.2, -0.3], [0.2, 0.2, 0.1], (3000, 3))
    cn = rng.normal(size=(3000, 3)); cn /= np.linalg.norm(cn, axis=1, keepdims=True)
    hb_pts = np.concatenate([sp, clutter]); hb_nrm = np.concatenate([bn[vis], cn])
    scene_xyz = (hb_pts @ hic_true[:3, :3].T + hic_true[:3, 3]).astype(np.float32)
    scene_nrm = (hb_nrm @ hic_true[:3, :3].T).astype(np.float32)
    f11 = np.eye(4); f11[:3, 3] = [0.0, -0.045, 0.02]
    f21 = np.eye(4); f21[:3, 3] = [0.0, 0.045, 0.02]
    scene, base = ctx.upload_cloud(scene_xyz, scene_nrm), ctx.upload_cloud(base_xyz, base_nrm)
    new_hic, matched, offset = hand.handbase_icp(ctx, scene, base, hic_hat.astype(np.float32), f11, f21)
    # the restated chain on the device's voxel grid output
    ds = scene.voxel_grid(0.005)
    dx, dn, _ = ds.download()
    hx, hn = O.transform_cloud(np.linalg.inv(hic_hat.astype(np.float32)), dx, dn)
    m = (hx[:, 0] >= np.float32(-0.07)) & (hx[:, 0] <= np.float32(0.03))
    hx, hn = hx[m], hn[m]
    m = (hx[:, 2] >= np.float32(-0.18)) & (hx[:, 2] <= np.float32(0.01))
    hx, hn = hx[m], hn[m]
    y1, z1, y2, z2 = np.float32(-0.045), np.float32(0.02), np.float32(0.045), np.float32(0.02)
    d1 = ((hx[:, 2] - z1) ** 2 + (hx[:, 1] - y1) ** 2).astype(np.float64) <= 0.015 * 0.015
    d2 = ((hx[:, 2] - z2) ** 2 + (hx[:, 1] - y2) ** 2).astype(np.float64) <= 0.015 * 0.015
    strip = (hx[:, 1] >= y1) & (hx[:, 1] <= y2) & (np.abs(hx[:, 2] - z1).astype(np.float64) <= 0.01)
    keep = ~(d1 | d2 | strip)
    rx, rn = hx[keep], hn[keep]
    tmp = ctx.upload_cloud(dx, dn)
    t1 = tmp.transform(np.linalg.inv(hic_hat.astype(np.float32))); t2 = t1.pass_through("x", -0.07, 0.03); t3 = t2.pass_through("z", -0.18, 0.01)
    t4 = t3.handbase_region(float(y1), float(z1), float(y2), float(z2))
    gx, gn, _ = t4.download()
    assert np.array_equal(gx, rx) and np.array_equal(gn, rn) and len(rx) > 200
    T_ref, it, conv = O.run_icp(rx, rn, base_xyz, base_nrm, max_iter=50, angle=30.0, dist=0.03)
    dt, dr = synth.pose_error(offset[None], np.asarray(T_ref, np.float32).reshape(1, 4, 4))
    assert dt[0] <= 1e-3 and dr[0] <= 1.0, (dt, dr)
    assert matched
    et, er = synth.pose_error(new_hic[None], hic_true.astype(np.float32)[None])
    e0, r0 = synth.pose_error(hic_hat.astype(np.float32)[None], hic_true.astype(np.float32)[None])
    assert et[0] < 0.003 and er[0] < 1.5 and et[0] < e0[0]
    for c in (scene, base, ds, tmp, t1, t2, t3, t4):
        c.free()
