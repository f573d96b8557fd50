"""The device versions of PCL's two normal estimators (hop_frame_organized, hop_cloud_mls) against their oracle restatements:
positions / valid sets bit-exact, integral-image normals bit-exact (same double integral image, same operation order), MLS to 1e-5
(double sums over the neighbours in a different order)."""
import os

import numpy as np
import pytest

from hop_b200 import synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu
K = (616.5961303710938, 616.59619140625, 307.6278076171875, 239.68692016601562)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _frames():
    rng = np.random.default_rng(3)
    # (1) a synthetic 640 x 480 frame: an ellipsoid in front of a tilted wall, millimetre quantisation + noise, holes
    dense, _ = synth.make_model("ellipse", 300000, seed=7)
    T = np.eye(4); T[:3, :3] = synth._rot_from_rotvec(np.array([0.4, -0.7, 0.3])); T[:3, 3] = [0.0, 0.005, 0.35]
    P = dense.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    u = np.round(P[:, 0] * K[0] / P[:, 2] + K[2]).astype(int); v = np.round(P[:, 1] * K[1] / P[:, 2] + K[3]).astype(int)
    vv, uu = np.meshgrid(np.arange(640), np.arange(480))
    wall = 0.55 + 0.1 * (vv - 320) / 640.0
    img = wall.copy(); np.minimum.at(img, (v, u), P[:, 2])
    depth = np.round(img * 1000 + rng.normal(0, 0.6, img.shape)).astype(np.uint16)
    depth[rng.random(depth.shape) < 0.01] = 0
    depth[100:140, 500:560] = 0
    yield "synthetic", depth
    # (2) the frame the reference ships (example/depth7.png, kept as an input fixture)
    p = os.path.join(GOLD, "example_depth7.png")
    if os.path.exists(p):
        import cv2
        d = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        if d is not None and d.dtype == np.uint16:
            yield "example_depth7", d


def test_frame_organized_matches_oracle(ctx):
    for name, depth in _frames():
        fp = ctx.frame_params(K=K)
        cloud = ctx.frame_organized(depth, fp)
        xyz, nrm, conf = cloud.download()
        ref_xyz = O.organized_cloud(depth, K)
        ref_nrm = O.integral_image_normals(ref_xyz)
        keep = (ref_xyz[..., 2].astype(np.float64) >= 0.1) & (ref_xyz[..., 2].astype(np.float64) <= 2.0)
        assert len(xyz) == keep.sum(), name
        assert np.array_equal(xyz, ref_xyz[keep]), name                     # raster order, bit-exact back-projection
        rn = ref_nrm[keep]
        assert np.array_equal(np.isnan(nrm[:, 0]), np.isnan(rn[:, 0])), name  # the same pixels carry a normal
        ok = ~np.isnan(rn[:, 0])
        assert ok.mean() > 0.5, name
        assert np.array_equal(nrm[ok], rn[ok]), (name, np.abs(nrm[ok] - rn[ok]).max())
        cloud.free()


def test_cloud_mls_matches_oracle(ctx):
    rng = np.random.default_rng(1)
    m, mn = synth.make_model("ellipse", 6000, seed=2)
    pts = (m[m[:, 2] > -0.005] + rng.normal(0, 2e-4, (np.sum(m[:, 2] > -0.005), 3)) + [0.01, -0.02, 0.35]).astype(np.float32)
    pts = np.concatenate([pts, [[0.3, 0.3, 0.9], [0.31, 0.3, 0.9]]]).astype(np.float32)   # two points with < 3 neighbours
    conf = rng.random(len(pts)).astype(np.float32)
    cloud = ctx.upload_cloud(pts, None, conf)
    out = cloud.mls(0.003)
    xyz, nrm, c = out.download()
    po, no, valid = O.mls(pts, 0.003)
    assert len(xyz) == valid.sum() and not valid[-1] and not valid[-2]
    assert np.array_equal(c, conf[valid])                                  # the confidence travels with the kept points, in order
    assert np.abs(xyz - po[valid]).max() < 2e-7
    cosang = np.einsum("ij,ij->i", nrm, no[valid])
    assert np.all(cosang > 1 - 1e-9 * 1e4), cosang.min()                   # same orientation too (pcl::eigen33's sign), to 1e-5
    assert np.abs(nrm - no[valid]).max() < 1e-4
    cloud.free(); out.free()
