"""The restatements of PCL's two normal estimators (oracle/hop_oracle_frame.c; PCL is not installed: parity unpinned against PCL)
against closed forms: a plane and a sphere seen by the shipped camera, depth discontinuities, the image border."""
import numpy as np

from oracle import cpu_oracle as O

K = (616.5961303710938, 616.59619140625, 307.6278076171875, 239.68692016601562)


def _render_plane(normal, d0, w=160, h=120, K=K):
    """depth image [mm] of the plane n . X = d0 (n pointing away from the camera), quantised to millimetres"""
    fx, fy, cx, cy = K
    v, u = np.meshgrid(np.arange(w), np.arange(h))
    ray = np.stack([(v - cx) / fx, (u - cy) / fy, np.ones_like(v, float)], -1)
    z = d0 / (ray @ np.asarray(normal, float))
    return np.round(z * 1000).astype(np.uint16)


def test_integral_image_normals_of_a_plane():
    n_true = np.array([0.2, -0.1, 1.0]); n_true /= np.linalg.norm(n_true)
    Ks = (K[0] / 4, K[1] / 4, 80.0, 60.0)
    depth = _render_plane(n_true, 0.5, K=Ks)
    xyz = O.organized_cloud(depth, Ks)
    nrm = O.integral_image_normals(xyz)
    inner = nrm[10:-10, 10:-10]
    assert np.isfinite(inner).all()
    # flipped towards the camera: n . (0 - p) > 0, i.e. the opposite of the plane's away-pointing normal
    cosang = np.abs(inner @ n_true)
    assert cosang.min() > np.cos(np.radians(3.0)) and np.median(cosang) > np.cos(np.radians(0.5))
    assert np.all(np.einsum("ijk,ijk->ij", inner, -xyz[10:-10, 10:-10]) > 0)
    # border policy IGNORE: the outer `smoothing` = 10 pixels carry no normal
    assert np.isnan(nrm[:10]).all() and np.isnan(nrm[-10:]).all() and np.isnan(nrm[:, :10]).all() and np.isnan(nrm[:, -10:]).all()


def test_integral_image_normals_stop_at_depth_discontinuities_and_invalid_pixels():
    Ks = (K[0] / 4, K[1] / 4, 80.0, 60.0)
    depth = _render_plane([0, 0, 1.0], 0.5, K=Ks)
    depth[:, 80:] = _render_plane([0, 0, 1.0], 0.9, K=Ks)[:, 80:]      # a 40 cm step at column 80
    depth[40:50, 30:40] = 0                                             # a hole (invalid pixels are (0, 0, 0) in the reference)
    xyz = O.organized_cloud(depth, Ks)
    nrm = O.integral_image_normals(xyz)
    # within 2 pixels of the step the smoothing rectangle collapses (distance <= 2): no normal
    assert np.isnan(nrm[20:100, 79:82]).all()
    # a few pixels away the rectangle is small but the normal is that of the fronto-parallel plane, towards the camera
    ok = nrm[20:100, 60:70]
    assert np.isfinite(ok).all() and np.all(ok[..., 2] < -0.99)
    assert np.isnan(nrm[40:50, 30:40]).all() or np.all(xyz[40:50, 30:40] == 0)
    far = nrm[20:100, 100:140]
    assert np.isfinite(far).all() and np.all(far[..., 2] < -0.99)


def test_mls_on_a_sphere_projects_and_gives_radial_normals():
    rng = np.random.default_rng(0)
    r = 0.03
    d = rng.normal(size=(6000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d[d[:, 2] < -0.3]                                              # a cap, like a visible surface
    pts = (d * r + np.array([0, 0, 0.4]) + rng.normal(0, 2e-4, d.shape)).astype(np.float32)
    po, no, valid = O.mls(pts, 0.003)
    assert valid.mean() > 0.98
    c = np.array([0, 0, 0.4])
    rad = (po[valid] - c); rn = np.linalg.norm(rad, axis=1)
    # projection pulls the noisy points back onto the sphere: radial error shrinks (0.55 x with ~15 neighbours per point)
    raw = np.abs(np.linalg.norm(pts[valid] - c, axis=1) - r)
    assert np.abs(rn - r).mean() < 0.7 * raw.mean()
    cosang = np.abs(np.einsum("ij,ij->i", no[valid], rad / rn[:, None]))
    # (0.2 mm of noise over a 3 mm neighbourhood: a few degrees of normal noise)
    assert np.median(cosang) > np.cos(np.radians(5.0)) and np.percentile(cosang, 5) > np.cos(np.radians(25.0))
    assert np.allclose(np.linalg.norm(no[valid], axis=1), 1.0, atol=1e-5)
    # fewer than three neighbours: dropped from the corresponding indices
    lonely = np.concatenate([pts[:50], [[1, 1, 1.0]]]).astype(np.float32)
    _, _, v2 = O.mls(lonely, 0.003)
    assert not v2[-1]
