"""The C++ host side (icra20-hand-object-pose_b200/host): YAML-subset reader vs PyYAML, PNG16 reader vs cv2, VoxelGrid /
depth back-projection vs numpy restatements of the reference's Utils (CPU), and main_realdata_auto end to end on a
synthetic depth frame (GPU)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from hop_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "icra20-hand-object-pose_b200", "host")
TOOL = os.path.join(HOST, "host_tool")
MAIN = os.path.join(HOST, "main_realdata_auto")

CONFIG = """##Motoman left arm realsense
cam_K: [616.5961303710938, 0.0, 307.6278076171875, 0.0, 616.59619140625, 239.68692016601562, 0.0, 0.0, 1.0]

# K used in blender
# cam_K: [619.2578, 0.0, 320,
#         0.0, 619.2578, 240]

# camera xyz, q(xyzw)
cam1_in_leftarm: [0.0,0.0,0.0,0.0,0.0,0.0,1.0]

handbase_in_palm: [1    ,               0             ,       0 ,0,
                  -0              ,      1          ,          0  ,0,
                  0     ,              -0       ,             1 , 0,
                  -0         ,          -0         ,          -0      ,              1
]

out_dir: {out}
rgb_path: {out}/rgb.png
depth_path: {out}/depth.png
palm_in_baselink: {out}/palm_in_base.txt
leftarm_in_base: {out}/arm_left.txt
model_name: ellipse
object_model_path: {out}/ellipse.ply
object_mesh_path: {out}/ellipse.obj
ppf_path: {out}/ppf_ellipse

object_symmetry:   # 0 means complete symmetric, 360 means not symmetric
  tless3:
    x: 360
    y: 0
    z: 360
  ellipse:
    x: 180
    y: 180
    z: 180

down_sample:
  leaf_size: 0.005
remove_noise:
  radius: 0.01
  min_number: 10
hand_match:
  finger1_min_match: 5
  finger2_dist_thres: 0.005
  check_normal: true
  pso:
    n_pop: 15
    n_gen: 3   #Increase for better accuracy but slower
    pso_par_initial_w: 0.0      # Particle velocity initial scaling factor
lcp:
  dist: 0.001
  normal_angle: 10

pose_estimator_wrong_ratio: 1
pose_estimator_high_confidence_thres: 0.8
icp_dist_thres: 0.01
icp_angle_thres: 45
super4pcs_sample_size: 100    #sample points on Q(model)
super4pcs_overlap: 0.2
super4pcs_delta: 0.003
super4pcs_dispersion: 0.5
super4pcs_success_quadrilaterals: 10   #Increase for better accuracy but slower
super4pcs_max_normal_difference: -1    #Maximum angle (degrees) between corresponded normals.
super4pcs_max_color_distance: -1
super4pcs_max_time_seconds: 1

urdf_path: {out}/hand_T42b.urdf
Hand:
  base_link:
    mesh: {out}/base.obj
    cloud: {out}/base.ply
"""


def _tool(*args):
    r = subprocess.run([TOOL, *map(str, args)], capture_output=True, text=True)
    return r.returncode, r.stdout


def _write_ply(path, xyz, nrm, binary=True):
    n = len(xyz)
    hdr = f"ply\nformat {'binary_little_endian' if binary else 'ascii'} 1.0\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\n" \
          "property float nx\nproperty float ny\nproperty float nz\nend_header\n"
    with open(path, "wb") as f:
        f.write(hdr.encode())
        data = np.concatenate([xyz, nrm], 1).astype("<f4")
        if binary:
            f.write(data.tobytes())
        else:
            f.write("\n".join(" ".join(repr(float(v)) for v in row) for row in data).encode() + b"\n")


def _read_ply_ascii(path):
    lines = open(path).read().split("\n")
    k = lines.index("end_header")
    return np.array([[float(v) for v in l.split()] for l in lines[k + 1:] if l.strip()], np.float64)


needs_tool = pytest.mark.skipif(not os.path.exists(TOOL), reason="host tools not built (make -C icra20-hand-object-pose_b200/host)")


@needs_tool
def test_yaml_subset_reader_matches_pyyaml(tmp_path):
    import yaml
    path = tmp_path / "cfg.yaml"
    path.write_text(CONFIG.format(out=str(tmp_path)))
    ref = yaml.safe_load(path.read_text())
    for key in ["out_dir", "model_name", "lcp.dist", "lcp.normal_angle", "hand_match.pso.n_pop", "hand_match.pso.pso_par_initial_w",
                "hand_match.check_normal", "object_symmetry.ellipse.z", "object_symmetry.tless3.y", "super4pcs_delta",
                "super4pcs_max_normal_difference", "Hand.base_link.cloud", "urdf_path", "down_sample.leaf_size"]:
        rc, out = _tool("yaml", path, key)
        v = ref
        for part in key.split("."):
            v = v[part]
        assert rc == 0
        if isinstance(v, bool):
            assert out.strip() == str(v).lower()
        elif isinstance(v, (int, float)):
            assert float(out) == float(v), key
        else:
            assert out.strip() == str(v), key
    for key in ["cam_K", "cam1_in_leftarm", "handbase_in_palm"]:          # flow sequences, one of them over five lines
        rc, out = _tool("yaml", path, key)
        assert rc == 0 and np.allclose([float(x) for x in out.split()], ref[key])
    rc, out = _tool("yaml", path, "object_symmetry")
    assert out.split() == list(ref["object_symmetry"].keys())
    assert _tool("yaml", path, "no_such_key")[0] == 3
    cfg = os.path.join("/root/reference", "config_autodataset.yaml")      # the reference's own file, when it is around
    if os.path.exists(cfg):
        full = yaml.safe_load(open(cfg).read())
        rc, out = _tool("yaml", cfg, "hand_match.outter_pt_dist_weight")
        assert rc == 0 and float(out) == float(full["hand_match"]["outter_pt_dist_weight"])
        rc, out = _tool("yaml", cfg, "Hand.finger_2_2.convex_mesh")
        assert out.strip() == full["Hand"]["finger_2_2"]["convex_mesh"]
        rc, out = _tool("yaml", cfg, "handbase_in_palm")
        assert np.allclose([float(x) for x in out.split()], full["handbase_in_palm"])


@needs_tool
def test_png16_reader_matches_cv2(tmp_path):
    import cv2
    rng = np.random.default_rng(0)
    smooth = (np.add.outer(np.arange(97), np.arange(131)) * 37 % 65536).astype(np.uint16)      # exercises the sub/up/paeth filters
    noisy = rng.integers(0, 65536, (64, 50)).astype(np.uint16)
    for k, img in enumerate([smooth, noisy]):
        p = str(tmp_path / f"d{k}.png")
        cv2.imwrite(p, img)
        rc, out = _tool("png", p)
        w, h, s, mn, mx = [int(x) for x in out.split()]
        assert rc == 0 and (w, h) == (img.shape[1], img.shape[0]) and s == int(img.astype(np.int64).sum()) and (mn, mx) == (img.min(), img.max())
    ex = "/root/reference/example/depth7.png"
    if os.path.exists(ex):
        img = cv2.imread(ex, cv2.IMREAD_UNCHANGED)
        rc, out = _tool("png", ex)
        assert [int(x) for x in out.split()] == [img.shape[1], img.shape[0], int(img.astype(np.int64).sum()), int(img.min()), int(img.max())]


def _voxel_numpy(xyz, nrm, leaf):
    """pcl::VoxelGrid restated: floor(coord * (1/leaf)), leaves ordered by linear index (x fastest), centroid per leaf."""
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(xyz.astype(np.float32) * inv).astype(np.int64)
    mn = np.floor(xyz.min(0).astype(np.float32) * inv).astype(np.int64)
    mx = np.floor(xyz.max(0).astype(np.float32) * inv).astype(np.int64)
    div = mx - mn + 1
    idx = (ijk[:, 0] - mn[0]) + (ijk[:, 1] - mn[1]) * div[0] + (ijk[:, 2] - mn[2]) * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    uniq, start = np.unique(idx[order], return_index=True)
    out_p, out_n = [], []
    for a, b in zip(start, list(start[1:]) + [len(order)]):
        sel = order[a:b]
        out_p.append(xyz[sel].astype(np.float64).mean(0))
        n = nrm[sel].astype(np.float64).sum(0)
        out_n.append(n / max(np.linalg.norm(n), 1e-30))
    return np.array(out_p), np.array(out_n)


@needs_tool
@pytest.mark.parametrize("binary", [True, False])
def test_voxel_grid_and_ply_io(tmp_path, binary):
    m, mn = synth.make_model("cuboid", 5000, seed=3)
    src, dst = str(tmp_path / "in.ply"), str(tmp_path / "out.ply")
    _write_ply(src, m, mn, binary)
    assert _tool("voxel", src, 0.005, dst)[0] == 0
    got = _read_ply_ascii(dst)
    ref_p, ref_n = _voxel_numpy(m, mn, 0.005)
    assert len(got) == len(ref_p) and 200 < len(got) < 5000
    assert np.abs(got[:, :3] - ref_p).max() < 1e-6 and np.abs(got[:, 3:6] - ref_n).max() < 1e-5


URDF = """<?xml version="1.0"?>
<!-- a T42-like hand: the subset Hand::parseURDF reads -->
<robot name="hand">
  <link name="base_link">
    <visual><origin xyz="0 0 0" rpy="0 0 0"/><geometry><mesh filename="package://x/base.STL" scale="0.001 0.001 0.001"/></geometry></visual>
  </link>
  <link name="rail_1"><visual><geometry><box size="1 1 1"/></geometry></visual></link>
  <link name="finger_1_1">
    <visual>
      <origin xyz="0.01 -0.02 0.03" rpy="0.1 -0.2 1.5707963"/>
      <geometry><mesh filename='f11.STL' scale="1 2 3"/></geometry>
    </visual>
    <collision><geometry><box size="1 1 1"/></geometry></collision>
  </link>
  <link name="finger_1_2"><visual><geometry><mesh filename="f12.STL"/></geometry></visual></link>
  <joint name="j1" type="revolute">
    <parent link="base_link"/> <child link="finger_1_1"/>
    <origin xyz="-0.15 -0.04 0.02" rpy="3.1415926 0 0.5"/> <axis xyz="1 0 0"/>
  </joint>
  <joint name="j2" type="revolute"><origin rpy="0 0.3 0" xyz="0 0 -0.06"/><parent link="finger_1_1"/><child link="finger_1_2"/></joint>
</robot>
"""


URDF_TRICKY = """<?xml version="1.0" encoding="utf-8"?>
<!DOCTYPE robot>
<!-- <link name="ghost"/> inside a comment -->
<robot name = "h" xmlns:xacro="http://www.ros.org/wiki/xacro">
  <material name="grey"><color rgba="0.5 0.5 0.5 1"/></material>
  <link name="base_link">
    <inertial><mass value="1"/><origin xyz="9 9 9"/></inertial>
    <visual>
      <origin rpy = '0 0 0'   xyz="1e-3 -2E-3 .5"/>
      <geometry><mesh filename="package://a/b&gt;c.STL" scale="0.001 0.001 0.001" /></geometry>
      <material name="grey"/>
    </visual>
  </link>
  <link name="finger_1_1"><visual><geometry><mesh filename="a>b.STL"/></geometry></visual>text &amp; more</link>
  <joint name="j" type="revolute"><origin xyz="0 1 2"
      rpy="0 0 1.5707963267948966"/><parent link="base_link"/><child link="finger_1_1"/><limit lower="0" upper="1"/></joint>
</robot>
"""


@needs_tool
@pytest.mark.parametrize("URDF,names", [(URDF, ["base_link", "finger_1_1", "finger_1_2"]), (URDF_TRICKY, ["base_link", "finger_1_1"])],
                         ids=["t42_like", "declarations_comments_quotes_entities"])
def test_urdf_reader_matches_elementtree(tmp_path, URDF, names):
    """the built-in XML-subset reader + parseUrdfLinks (what Hand::parseURDF reads, Hand.cpp:375-502) against xml.etree + scipy"""
    import xml.etree.ElementTree as ET
    from scipy.spatial.transform import Rotation
    (tmp_path / "hand.urdf").write_text(URDF)
    rc, out = _tool("urdf", tmp_path / "hand.urdf")
    assert rc == 0, out
    rows = [l.split() for l in out.strip().split("\n")]
    assert [r[0] for r in rows] == names                                               # the rail is skipped
    root = ET.fromstring(URDF)

    def pose(el):
        T = np.eye(4)
        if el is not None:
            rpy = [float(v) for v in el.get("rpy", "0 0 0").split()]
            T[:3, :3] = Rotation.from_euler("ZYX", [rpy[2], rpy[1], rpy[0]]).as_matrix()   # Rz(yaw) Ry(pitch) Rx(roll)
            T[:3, 3] = [float(v) for v in el.get("xyz", "0 0 0").split()]
        return T
    for r in rows:
        name, parent = r[0], r[1]
        vals = np.array([float(v) for v in r[2:]])
        link = [l for l in root.findall("link") if l.get("name") == name][0]
        joint = [j for j in root.findall("joint") if j.find("child").get("link") == name]
        assert parent == (joint[0].find("parent").get("link") if joint else "-")
        mesh = link.find("visual/geometry/mesh")
        scale = [float(v) for v in mesh.get("scale", "1 1 1").split()]
        assert np.allclose(vals[:3], scale)
        assert np.abs(vals[3:19].reshape(4, 4) - pose(link.find("visual/origin"))).max() < 1e-6
        assert np.abs(vals[19:35].reshape(4, 4) - pose(joint[0].find("origin") if joint else None)).max() < 1e-6
    for bad in ("<robot><link name='a'></robot>", "<robot><link name=a/></robot>", "<robot>", "no xml at all"):
        (tmp_path / "bad.urdf").write_text(bad)
        assert _tool("urdf", tmp_path / "bad.urdf")[0] == 3
    assert _tool("urdf", tmp_path / "missing.urdf")[0] == 3


@needs_tool
def test_obj_mesh_reader(tmp_path):
    """loadOBJMesh (igl::readOBJ stand-in of SDFchecker::registerMesh): plain, v/vt/vn and negative indices, quads fanned"""
    V, F = synth.make_mesh("tless", 2)
    vol = np.einsum("ij,ij->i", V[F[:, 0]].astype(float), np.cross(V[F[:, 1]].astype(float), V[F[:, 2]].astype(float))).sum() / 6
    vs = "".join("v %.9g %.9g %.9g\n" % tuple(v) for v in V)
    (tmp_path / "a.obj").write_text("# comment\n" + vs + "vn 0 0 1\n" + "".join("f %d %d %d\n" % tuple(f + 1) for f in F))
    (tmp_path / "b.obj").write_text(vs + "".join("f %d/1/1 %d//1 %d\n" % (f[0] + 1, f[1] + 1, f[2] - len(V)) for f in F))
    for name in ("a.obj", "b.obj"):
        rc, out = _tool("obj", tmp_path / name)
        assert rc == 0, out
        nv, nf, v, f0, f1, f2 = out.split()
        assert (int(nv), int(nf)) == (len(V), len(F)) and abs(float(v) - vol) < 1e-9 and [int(f0), int(f1), int(f2)] == list(F[0])
    (tmp_path / "q.obj").write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n")
    rc, out = _tool("obj", tmp_path / "q.obj")
    assert rc == 0 and out.split()[:2] == ["4", "2"]
    (tmp_path / "bad.obj").write_text("v 0 0 0\nv 1 0 0\nf 1 2 9\n")
    assert _tool("obj", tmp_path / "bad.obj")[0] == 3
    assert _tool("obj", tmp_path / "missing.obj")[0] == 3


@needs_tool
def test_depth_backprojection(tmp_path):
    import cv2
    rng = np.random.default_rng(1)
    d = rng.integers(0, 2500, (48, 64)).astype(np.uint16)          # millimetres; < 100 and > 2000 are invalid (Utils.cpp:45)
    p = str(tmp_path / "depth.png")
    cv2.imwrite(p, d)
    fx, fy, cx, cy = 616.59613, 616.59619, 30.5, 20.25
    out = str(tmp_path / "cloud.ply")
    assert _tool("depthcloud", p, fx, fy, cx, cy, out)[0] == 0
    got = _read_ply_ascii(out)[:, :3]
    z = d.astype(np.float32) * np.float32(0.001)
    valid = ~((z > 2.0) | (z < 0.1)) & (z > 0.1) & (z < 2.0)
    u, v = np.nonzero(valid)
    ref = np.stack([(v - np.float32(cx)) * z[u, v] / np.float32(fx), (u - np.float32(cy)) * z[u, v] / np.float32(fy), z[u, v]], 1)
    assert len(got) == len(ref) and np.abs(got - ref).max() < 1e-6


@needs_tool
def test_normals_of_a_plane_face_the_viewpoint(tmp_path):
    rng = np.random.default_rng(2)
    xy = rng.uniform(-0.02, 0.02, (3000, 2))
    pts = np.concatenate([xy, np.full((3000, 1), 0.4)], 1).astype(np.float32)
    src, dst = str(tmp_path / "plane.ply"), str(tmp_path / "plane_n.ply")
    _write_ply(src, pts, np.zeros_like(pts))
    assert _tool("normals", src, 0.003, dst)[0] == 0
    n = _read_ply_ascii(dst)[:, 3:6]
    ok = np.isfinite(n).all(1)
    assert ok.mean() > 0.95 and np.allclose(n[ok], [0, 0, -1], atol=1e-4)


# ---------------------------------------------------------------------------------------------------------------- GPU
def _render_depth(model_xyz, pose, K, shape=(480, 640)):
    P = model_xyz @ pose[:3, :3].T + pose[:3, 3]
    u = np.round(P[:, 0] * K[0] / P[:, 2] + K[2]).astype(int)
    v = np.round(P[:, 1] * K[1] / P[:, 2] + K[3]).astype(int)
    ok = (u >= 0) & (u < shape[1]) & (v >= 0) & (v < shape[0])
    depth = np.full(shape, np.inf)
    np.minimum.at(depth, (v[ok], u[ok]), P[ok, 2])
    depth[~np.isfinite(depth)] = 0
    return np.round(depth * 1000).astype(np.uint16)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(MAIN), reason="main_realdata_auto not built")
def test_main_realdata_auto_end_to_end(tmp_path):
    """main_realdata_auto <config.yaml> on a synthetic depth frame of the ellipsoid: the written model2scene.txt must put the
    model within 2 mm (ADI: mean closest-point distance, the reference's own metric, scripts/eval_utils.py:181-200)."""
    import cv2
    from scipy.spatial import cKDTree
    out = str(tmp_path)
    (tmp_path / "cfg.yaml").write_text(CONFIG.format(out=out))
    np.savetxt(out + "/arm_left.txt", np.eye(4))
    hb = np.eye(4); hb[:3, 3] = [0.15, 0.0, 0.38]                     # hand base in the camera frame
    np.savetxt(out + "/palm_in_base.txt", hb)
    dense, dn = synth.make_model("ellipse", 400000, seed=7)
    model, mn = synth.make_model("ellipse", 60000, seed=8)
    _write_ply(out + "/ellipse.ply", model, mn, binary=True)
    mV, mF = synth.make_mesh("ellipse", 3)                             # object_mesh_path: the physics pruning step registers it
    with open(out + "/ellipse.obj", "w") as f:
        f.write("".join("v %.9g %.9g %.9g\n" % tuple(v) for v in mV) + "".join("f %d//%d %d//%d %d//%d\n" % (a, a, b, b, c, c) for a, b, c in mF + 1))
    gt = np.eye(4)
    gt[:3, :3] = synth._rot_from_rotvec(np.array([0.4, -0.7, 0.3]))
    gt[:3, 3] = [0.0, 0.005, 0.35]                                     # = (-0.15, 0.005, -0.03) in the hand-base frame: inside the crop box
    K = (616.5961303710938, 616.59619140625, 307.6278076171875, 239.68692016601562)
    cv2.imwrite(out + "/depth.png", _render_depth(dense.astype(np.float64), gt, K))
    r = subprocess.run([MAIN, out + "/cfg.yaml"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "best tf:" in r.stdout and "num of pose clusters" in r.stdout
    assert "physics pruning:" in r.stdout                              # rejectByCollisionOrNonTouching ran (no hand model: first test only)
    assert "after projection check" in r.stdout                       # rejectByRender ran on the software rasteriser
    est = np.loadtxt(out + "/model2scene.txt")
    assert est.shape == (4, 4) and os.path.exists(out + "/best.obj") and os.path.exists(out + "/scene_normals.ply")
    sub = model[::20].astype(np.float64)
    a = sub @ est[:3, :3].T + est[:3, 3]
    b = sub @ gt[:3, :3].T + gt[:3, 3]
    adi = cKDTree(b).query(a)[0].mean()
    assert adi < 0.002, (adi, r.stdout[-1500:])


# ------------------------------------------------------------------------------------- the reference's shipped example frame
EXAMPLE_CFG = """cam_K: [616.5961303710938, 0.0, 307.6278076171875, 0.0, 616.59619140625, 239.68692016601562, 0.0, 0.0, 1.0]
cam1_in_leftarm: [-0.004269333556294441,-0.007711530197411776,-0.08680825680494308,-0.006834600586444139,0.9986741542816162,0.04945759475231171,-0.01254322659224272]
handbase_in_palm: [0    ,               -1             ,       0 ,0.009000016543892384,
                  -0              ,      0          ,          1  ,0.08699987083673477,
                  -1     ,              -0       ,             0 , 0.01899999752640724,
                  -0         ,          -0         ,          -0      ,              1
]
out_dir: {out}
depth_path: {gold}/example_depth7.png
palm_in_baselink: {gold}/example_palm_in_base7.txt
leftarm_in_base: {gold}/example_arm_left_link_7_t_7.txt
model_name: ellipse
object_model_path: {out}/ellipse.ply
"""


def example_frame_setup(tmp_path):
    """config of the shipped example frame (config_autodataset.yaml:2,12,14-18 + example/*) -> (depth uint16, K, cam_in_handbase)"""
    import cv2
    gold = os.path.join(os.path.dirname(__file__), "golden")
    (tmp_path / "ex.yaml").write_text(EXAMPLE_CFG.format(out=str(tmp_path), gold=gold))
    rc, out = _tool("config", str(tmp_path / "ex.yaml"))
    assert rc == 0, out
    lines = out.strip().split("\n")
    k = lines.index("handbase_in_cam")
    hic = np.array([[float(v) for v in l.split()] for l in lines[k + 1:k + 5]])
    K = [float(v) for v in lines[-1].split()[1:]]
    depth = cv2.imread(os.path.join(gold, "example_depth7.png"), cv2.IMREAD_UNCHANGED)
    return depth, K, np.linalg.inv(hic).astype(np.float32)


@needs_tool
def test_example_frame_plumbing(tmp_path):
    """BASELINE config C1 (plumbing, no GPU): the reference's shipped depth frame through the host front end.  SURVEY quick
    facts: 480x640 uint16 mm, 68 600 valid pixels, median 0.349 m; the hand region must come out non-empty."""
    depth, K, cam_in_handbase = example_frame_setup(tmp_path)
    assert depth.shape == (480, 640) and depth.dtype == np.uint16
    d = depth.astype(np.float32) * np.float32(0.001)
    valid = (d > 0.1) & (d < 2.0)
    assert valid.sum() == 68600 and abs(float(np.median(d[valid])) - 0.349) < 0.002
    np.savetxt(tmp_path / "T.txt", cam_in_handbase)
    rc, out = _tool("frame", os.path.join(os.path.dirname(__file__), "golden", "example_depth7.png"), *K, str(tmp_path / "T.txt"), str(tmp_path / "seg.bin"))
    assert rc == 0, out
    raw = np.fromfile(tmp_path / "seg.bin", np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    seg = raw[4:4 + 28 * n].view(np.float32).reshape(n, 7)
    assert 200 < n < 20000                                   # the object + fingers inside the hand-base crop box, 3 mm voxels
    assert np.isfinite(seg).all() and np.allclose(np.linalg.norm(seg[:, 3:6], axis=1), 1.0, atol=1e-4)
    assert np.all(np.einsum("ij,ij->i", seg[:, :3], seg[:, 3:6]) <= 1e-6)   # normals face the camera
    hb = seg[:, :3] @ cam_in_handbase[:3, :3].T + cam_in_handbase[:3, 3]
    assert hb[:, 0].min() > -0.2505 and hb[:, 0].max() < -0.0695 and hb[:, 2].min() > -0.1205 and hb[:, 2].max() < 0.0505
