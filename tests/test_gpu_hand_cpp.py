"""The C++ Hand class (icra20-hand-object-pose_b200/host/Hand.{h,cpp}: setCurScene, matchOneComponentPSO x4, adjustHandHeight,
makeHandCloud, removeSurroundingPointsAndAssignProbability on the C ABI) driven by hand_demo the way main_realdata_auto.cpp:100-148
drives HandT42, on a synthetic two-finger hand (the reference's URDF and link clouds do not ship).  Every device stage has its own
parity test; this one checks the chain end to end against the ground truth of the synthetic frame: the four joint angles are
recovered, the hand base is left where it is, and the hand-point removal keeps the object and drops the hand."""
import os
import subprocess

import numpy as np
import pytest

from hop_b200 import synth
from test_host_cpp import _write_ply

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "icra20-hand-object-pose_b200", "host", "hand_demo")

CFG = """cam_K: [616.5961303710938, 0.0, 307.6278076171875, 0.0, 616.59619140625, 239.68692016601562, 0.0, 0.0, 1.0]
cam1_in_leftarm: [0.0,0.0,0.0,0.0,0.0,0.0,1.0]
handbase_in_palm: [1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1]
out_dir: {out}
depth_path: {out}/none.png
palm_in_baselink: {out}/eye.txt
leftarm_in_base: {out}/eye.txt
model_name: ellipse
object_model_path: {out}/none.ply
gripper_min_dist: 0.02
near_hand_dist: 0.003
hand_match:
  finger1_min_match: 5
  finger2_min_match: 5
  finger1_dist_thres: 0.005
  finger2_dist_thres: 0.005
  finger1_normal_angle: 60
  finger2_normal_angle: 60
  check_normal: true
  max_outter_pts: 300
  outter_pt_dist: 0.002
  outter_pt_dist_weight: 1
"""


def _rotx(deg):
    return synth._rot_x(np.deg2rad(deg))


def _hand(seed=1):
    rng = np.random.default_rng(seed)
    size = (0.02, 0.012, 0.06)
    links, tf_parent, parent = {}, {}, {}
    for name in ("finger_1_1", "finger_1_2", "finger_2_1", "finger_2_2"):
        links[name] = synth.make_finger_cloud(900, seed=seed + len(links), size=size)
    bp, bn = synth._cuboid(rng, 1500, 0.06, 0.13, 0.02)
    links["base_link"] = ((bp + [-0.09, 0.0, 0.035]).astype(np.float32), bn.astype(np.float32))
    left = np.eye(4); left[:3, 3] = [-0.15, -0.055, 0.02]
    right = np.eye(4); right[:3, :3] = np.diag([-1.0, -1.0, 1.0]); right[:3, 3] = [-0.15, 0.055, 0.02]
    out = np.eye(4); out[:3, 3] = [0, 0, -0.06]
    tf_parent.update(finger_1_1=left, finger_1_2=out, finger_2_1=right, finger_2_2=out, base_link=np.eye(4))
    parent.update(finger_1_1="base_link", finger_1_2="finger_1_1", finger_2_1="base_link", finger_2_2="finger_2_1", base_link="base_link")
    truth = dict(finger_1_1=10.0, finger_1_2=6.0, finger_2_1=8.0, finger_2_2=5.0)   # the tips stay > gripper_min_dist apart, around the object
    in_hb = {"base_link": np.eye(4)}
    for f in ("1", "2"):
        in_hb[f"finger_{f}_1"] = tf_parent[f"finger_{f}_1"] @ _rotx(truth[f"finger_{f}_1"])
        in_hb[f"finger_{f}_2"] = in_hb[f"finger_{f}_1"] @ out @ _rotx(truth[f"finger_{f}_2"])
    # the frame in the hand-base frame: the links where they really are (dense, noisy), the grasped object, a little clutter
    pts, nrm, is_hand = [], [], []
    for name, (x, n) in links.items():
        T = in_hb[name]
        dense = synth.make_finger_cloud(2500, seed=seed + 50 + len(pts), size=size) if name != "base_link" else (np.repeat(x, 2, 0), np.repeat(n, 2, 0))
        p = dense[0] + dense[1] * rng.normal(0, 0.0003, (len(dense[0]), 1))
        pts.append(p @ T[:3, :3].T + T[:3, 3]); nrm.append(dense[1] @ T[:3, :3].T); is_hand.append(np.ones(len(p), bool))
    op, on = synth._ellipsoid(rng, 3000, 0.022, 0.016, 0.015)
    pts.append(op + [-0.15, 0.0, -0.075]); nrm.append(on); is_hand.append(np.zeros(len(op), bool))
    hb_pts, hb_nrm, is_hand = np.concatenate(pts), np.concatenate(nrm), np.concatenate(is_hand)
    hic = np.eye(4); hic[:3, :3] = synth.random_rotation(rng); hic[:3, 3] = [0.05, -0.03, 0.45]
    cam_pts = (hb_pts @ hic[:3, :3].T + hic[:3, 3]).astype(np.float32)
    cam_nrm = (hb_nrm @ hic[:3, :3].T).astype(np.float32)
    return links, tf_parent, parent, truth, hic, cam_pts, cam_nrm, is_hand, hb_pts


def _urdf(links, tf_parent, parent, out):
    """the same hand as a URDF + Hand.<link>.cloud config entries (what Hand::parseURDF, Hand.cpp:375-502, reads): the link clouds are
    stored in millimetres with a 0.001 mesh scale, the joint poses as xyz / rpy (the right finger's diag(-1,-1,1) is a yaw of pi)"""
    xml, yml = ['<?xml version="1.0"?>', '<robot name="synthetic_hand">'], ["urdf_path: " + out + "/hand.urdf", "Hand:"]
    for name, (x, n) in links.items():
        _write_ply(out + f"/{name}_mm.ply", (x * 1000.0).astype(np.float32), n, binary=True)
        xml.append(f'  <link name="{name}"><visual><origin xyz="0 0 0" rpy="0 0 0"/><geometry><mesh filename="{name}.STL" scale="0.001 0.001 0.001"/></geometry></visual></link>')
        yml += [f"  {name}:", f"    cloud: {out}/{name}_mm.ply"]
    xml.append('  <link name="rail_1"><visual><geometry><mesh filename="rail.STL"/></geometry></visual></link>')
    for name, T in tf_parent.items():
        if name == "base_link":
            continue
        yaw = float(np.arctan2(T[1, 0], T[0, 0]))
        xml.append(f'  <joint name="j_{name}" type="revolute"><parent link="{parent[name]}"/><child link="{name}"/>'
                   f'<origin xyz="{float(T[0, 3])!r} {float(T[1, 3])!r} {float(T[2, 3])!r}" rpy="0 0 {yaw!r}"/><axis xyz="1 0 0"/></joint>')
    xml.append("</robot>")
    open(out + "/hand.urdf", "w").write("\n".join(xml) + "\n")
    return "\n".join(yml) + "\n"


@pytest.mark.skipif(not os.path.exists(DEMO), reason="hand_demo not built")
@pytest.mark.parametrize("source", ["links", "urdf"])
def test_cpp_hand_chain_recovers_the_grasp(tmp_path, source):
    links, tf_parent, parent, truth, hic, cam_pts, cam_nrm, is_hand, hb_pts = _hand()
    out = str(tmp_path)
    (tmp_path / "cfg.yaml").write_text(CFG.format(out=out) + (_urdf(links, tf_parent, parent, out) if source == "urdf" else ""))
    np.savetxt(out + "/eye.txt", np.eye(4))
    np.savetxt(out + "/handbase_in_cam.txt", hic)
    with open(out + "/links.txt", "w") as f:
        for name, (x, n) in links.items():
            _write_ply(out + f"/{name}.ply", x, n, binary=True)
            f.write(f"{name} {parent[name]} {name}.ply " + " ".join(repr(float(v)) for v in tf_parent[name].reshape(-1)) + "\n")
    _write_ply(out + "/scene_organized.ply", cam_pts, cam_nrm, binary=True)
    crop = (hb_pts[:, 0] >= -0.25) & (hb_pts[:, 0] <= -0.07) & (hb_pts[:, 2] >= -0.12) & (hb_pts[:, 2] <= 0.05)   # main_realdata_auto.cpp:79-93
    _write_ply(out + "/scene_hand_region.ply", cam_pts[crop], cam_nrm[crop], binary=True)
    r = subprocess.run([DEMO, out + "/cfg.yaml", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    lines = r.stdout.splitlines()
    matched = {l.split()[1]: int(l.split()[2]) for l in lines if l.startswith("match ")}
    assert matched == {"finger_1_1": 1, "finger_1_2": 1, "finger_2_1": 1, "finger_2_2": 1}, r.stdout[-2000:]
    for l in lines:
        if l.startswith("tf_self "):
            _, name, c, s = l.split()
            got = np.rad2deg(np.arctan2(float(s), float(c)))
            assert abs(got - truth[name]) < 3.0, (name, got, truth[name])
    k = lines.index("handbase_in_cam")
    hic_out = np.array([[float(v) for v in l.split()] for l in lines[k + 1:k + 5]])
    assert np.abs(hic_out - hic).max() < 6e-3                      # no spurious height correction beyond one 5 mm step
    last = [l for l in lines if l.startswith("hand cloud")][0].split()
    n_region, n_obj1 = int(last[4]), int(last[6])
    # what stays is the object (and stays confident), what goes is the hand
    import re
    txt = open(out + "/object1.ply").read().split("end_header\n")[1]
    kept = np.array([[float(v) for v in l.split()] for l in txt.strip().split("\n")])
    assert len(kept) == n_obj1 and 0 < n_obj1 < n_region
    from scipy.spatial import cKDTree
    _, idx = cKDTree(cam_pts[crop].astype(np.float64)).query(kept[:, :3])
    hand_kept = is_hand[crop][idx].mean()
    obj_total = (~is_hand[crop]).sum()
    obj_kept = (~is_hand[crop][idx]).sum()
    assert hand_kept < 0.35 and obj_kept > 0.6 * obj_total, (hand_kept, obj_kept, obj_total)
    conf = kept[:, 6]
    assert np.all((conf >= 0) & (conf <= 1)) and (conf[~is_hand[crop][idx]] > 0.8).mean() > 0.3
