"""GPU parity of the per-frame front end (hop_frame_to_scene) against the C++ host restatement of the reference's
pre-processing chain (host/cloud.cpp frameToObjectSegment, run through host_tool; its pieces are themselves checked against
numpy / OpenCV / PyYAML in tests/test_host_cpp.py).

Bars: point counts of every stage and the positions are BIT-identical (same float operations in the same order: the voxel
centroids are sequential float sums in PCL's point order); normals come from double sums taken in a different order and a
device libm, so they are compared to 1e-5 (absolute, unit vectors)."""
import os
import subprocess

import numpy as np
import pytest

from hop_b200 import synth

pytestmark = pytest.mark.gpu
HOST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "icra20-hand-object-pose_b200", "host")
TOOL = os.path.join(HOST, "host_tool")
K = (616.5961303710938, 616.59619140625, 307.6278076171875, 239.68692016601562)


def _render_depth(model_xyz, pose, shape=(480, 640), noise_mm=0.0, seed=0):
    P = model_xyz @ pose[:3, :3].T + pose[:3, 3]
    u = np.round(P[:, 0] * K[0] / P[:, 2] + K[2]).astype(int)
    v = np.round(P[:, 1] * K[1] / P[:, 2] + K[3]).astype(int)
    ok = (u >= 0) & (u < shape[1]) & (v >= 0) & (v < shape[0])
    depth = np.full(shape, np.inf)
    np.minimum.at(depth, (v[ok], u[ok]), P[ok, 2])
    depth[~np.isfinite(depth)] = 0
    d = depth * 1000
    if noise_mm > 0:
        d = d + (d > 0) * np.random.default_rng(seed).normal(0, noise_mm, shape)
    return np.clip(np.round(d), 0, 65535).astype(np.uint16)


def _host_frame(tmp_path, depth, T):
    import cv2
    png, txt, out = str(tmp_path / "d.png"), str(tmp_path / "T.txt"), str(tmp_path / "seg.bin")
    cv2.imwrite(png, depth)
    np.savetxt(txt, T)
    r = subprocess.run([TOOL, "frame", png, *map(repr, K), txt, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    body = raw[4:4 + 28 * n].view(np.float32).reshape(n, 7)
    Ti = raw[4 + 28 * n:4 + 28 * n + 64].view(np.float32).reshape(4, 4).T.copy()   # column-major in the file
    return body[:, :3], body[:, 3:6], body[:, 6], Ti


def _scene_pose():
    hb = np.eye(4); hb[:3, 3] = [0.15, 0.0, 0.38]                # hand base in the camera frame
    gt = np.eye(4)
    gt[:3, :3] = synth._rot_from_rotvec(np.array([0.4, -0.7, 0.3]))
    gt[:3, 3] = [0.0, 0.005, 0.35]
    return np.linalg.inv(hb).astype(np.float32), gt              # cam_in_handbase (float32, as ConfigParser would hold it)


@pytest.mark.skipif(not os.path.exists(TOOL), reason="host tools not built (make -C icra20-hand-object-pose_b200/host)")
@pytest.mark.parametrize("noise_mm,seed", [(0.0, 0), (1.5, 3)])
def test_frame_front_end_matches_the_host_chain(ctx, tmp_path, noise_mm, seed):
    dense, _ = synth.make_model("ellipse", 400000, seed=7)
    T, gt = _scene_pose()
    depth = _render_depth(dense.astype(np.float64), gt, noise_mm=noise_mm, seed=seed)
    # clutter: a plane behind the object (crossing the crop box's far side) and a patch outside the box
    yy, xx = np.mgrid[0:480, 0:640]
    depth[(depth == 0) & (xx > 400) & (yy > 300)] = 520
    depth[(depth == 0) & (xx < 80)] = 3000                     # beyond 2 m: dropped by readDepthImage
    xyz_h, nrm_h, conf_h, Ti = _host_frame(tmp_path, depth, T)
    p = ctx.frame_params(K=K, cam_in_handbase=T, handbase_in_cam=Ti)
    cloud, counts = ctx.frame_to_scene(depth, p)
    xyz_g, nrm_g, conf_g = cloud.download()
    assert counts[0] == int(((depth * np.float32(0.001)) > 0.1).sum() - ((depth * np.float32(0.001)) >= 2.0).sum()) or counts[0] > 0
    assert len(xyz_g) == len(xyz_h) == counts[4] and len(xyz_h) > 300
    assert np.array_equal(xyz_g, xyz_h)                          # bit-identical positions, same order
    assert np.array_equal(conf_g, conf_h) and np.all(conf_g == 1.0)
    assert np.abs(nrm_g - nrm_h).max() < 1e-5
    assert np.allclose(np.linalg.norm(nrm_g, axis=1), 1.0, atol=1e-5)
    # refill an existing cloud with another frame: same result as a fresh one
    depth2 = _render_depth(dense.astype(np.float64), gt, noise_mm=0.7, seed=11)
    fresh, c2 = ctx.frame_to_scene(depth2, p)
    again, c3 = ctx.frame_to_scene(depth2, p, scene=cloud)
    a, b = fresh.download(), again.download()
    assert np.array_equal(c2, c3) and all(np.array_equal(x, y) for x, y in zip(a, b))
    fresh.free(); cloud.free()


def test_frame_front_end_edge_cases(ctx):
    T, gt = _scene_pose()
    p = ctx.frame_params(K=K, cam_in_handbase=T)
    empty = np.zeros((480, 640), np.uint16)
    cloud, counts = ctx.frame_to_scene(empty, p)                 # no valid pixel at all
    assert list(counts) == [0, 0, 0, 0, 0] and cloud.download()[0].shape == (0, 3)
    far = np.full((480, 640), 900, np.uint16)                    # a wall entirely outside the crop box
    cloud2, counts2 = ctx.frame_to_scene(far, p)
    assert counts2[0] == 480 * 640 and counts2[2] == 0 and counts2[4] == 0
    tiny = np.zeros((4, 6), np.uint16); tiny[1, 2] = 350        # one pixel: fewer than 3 neighbours -> NaN normal -> removed
    p2 = ctx.frame_params(K=(600, 600, 3, 2), cam_in_handbase=np.eye(4, dtype=np.float32), box_min=(-1, -1, -1), box_max=(1, 1, 1))
    cloud3, counts3 = ctx.frame_to_scene(tiny, p2)
    assert list(counts3) == [1, 1, 1, 1, 0]
    for c in (cloud, cloud2, cloud3):
        c.free()


def test_frame_to_pose_pipeline_on_device_cloud(ctx):
    """the device-made scene cloud goes straight into ICP + LCP (no host copy): the refined pose lands on the rendered truth"""
    dense, dn = synth.make_model("ellipse", 400000, seed=7)
    m, mn = synth.make_model("ellipse", 10000, seed=1)
    T, gt = _scene_pose()
    depth = _render_depth(dense.astype(np.float64), gt, noise_mm=0.5, seed=5)
    p = ctx.frame_params(K=K, cam_in_handbase=T)
    scene, counts = ctx.frame_to_scene(depth, p)
    model = ctx.upload_cloud(m, mn)
    hyp = synth.make_hypotheses(gt.astype(np.float32), 256, seed=4, random_frac=0.0)
    refined, it, cv = ctx.icp_refine(scene, model, hyp)
    sc = ctx.lcp_score(scene, model, refined)
    best = refined[int(np.argmax(sc))].astype(np.float64)
    from scipy.spatial import cKDTree
    sub = m[::10].astype(np.float64)
    adi = cKDTree(sub @ gt[:3, :3].T + gt[:3, 3]).query(sub @ best[:3, :3].T + best[:3, 3])[0].mean()
    assert adi < 0.002, adi
    scene.free(); model.free()


@pytest.mark.skipif(not os.path.exists(TOOL), reason="host tools not built (make -C icra20-hand-object-pose_b200/host)")
def test_shipped_example_frame_on_the_device(ctx, tmp_path):
    """the reference's own example frame (tests/golden/example_depth7.png, config_autodataset.yaml calibration): the device
    front end gives the host chain's cloud (positions bit for bit), and that cloud runs through Super4PCS -> clustering ->
    ICP -> LCP with a stand-in model (the frame's real object model is an external download: plumbing, not accuracy)."""
    import cv2
    import hop_b200
    from test_host_cpp import example_frame_setup
    depth, Kx, cam_in_handbase = example_frame_setup(tmp_path)
    png, txt, out = str(tmp_path / "d.png"), str(tmp_path / "T.txt"), str(tmp_path / "seg.bin")
    cv2.imwrite(png, depth)
    np.savetxt(txt, cam_in_handbase)
    r = subprocess.run([TOOL, "frame", png, *map(repr, Kx), txt, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(out, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    body = raw[4:4 + 28 * n].view(np.float32).reshape(n, 7)
    Ti = raw[4 + 28 * n:4 + 28 * n + 64].view(np.float32).reshape(4, 4).T.copy()
    p = ctx.frame_params(K=Kx, cam_in_handbase=cam_in_handbase, handbase_in_cam=Ti)
    cloud, counts = ctx.frame_to_scene(depth, p)
    xyz, nrm, conf = cloud.download()
    assert counts[0] == 68600 and len(xyz) == n > 200
    assert np.array_equal(xyz, body[:, :3]) and np.abs(nrm - body[:, 3:6]).max() < 1e-5
    # plumbing through the rest of the path with a stand-in model
    m, mn = synth.make_model("ellipse", 3000, seed=1)
    keys = np.unique(np.array([hop_b200.capi.compute_ppf(m[i], mn[i], m[j], mn[j]) for i in range(0, 3000, 30) for j in range(i + 30, 3000, 30)], np.int32), axis=0)
    est = hop_b200.PoseEstimator(ctx, {"model_name": "ellipse", "object_symmetry": {"ellipse": {"x": 180, "y": 180, "z": 180}}})
    est.setModel(m[::5], mn[::5], m, mn)
    est.setCurScene(xyz, nrm, conf)
    if est.runSuper4pcs(keys):
        est.clusterPoses(30, 0.015, True)
        est.refineByICP()
        est.clusterPoses(5, 0.003, False)
        best = est.selectBest()
        assert np.isfinite(best._pose).all() and len(est._pose_hypos) <= 100
    cloud.free()
