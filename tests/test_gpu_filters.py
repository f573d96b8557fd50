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
