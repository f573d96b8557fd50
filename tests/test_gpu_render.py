"""GPU parity tests of the render-based rejection (hop_render_scene_create, hop_render_depth, hop_reject_by_render) through the
C ABI against oracle/hop_oracle_render.c.  Bar: BIT-EXACT -- coverage is decided in integer arithmetic, depth in unfused
double, the wrong ratio with the reference's own row-major float sums, on both sides operation for operation.
(The oracle itself is unpinned against the reference's OpenGL renderer: tests/test_render_oracle.py.)"""
import numpy as np
import pytest

from hop_b200 import synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _params(ctx, cam, **kw):
    return ctx.render_params(fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], width=cam["width"], height=cam["height"], **kw)


def _oparams(cam, **kw):
    return O.render_params(fx=cam["fx"], fy=cam["fy"], cx=cam["cx"], cy=cam["cy"], width=cam["width"], height=cam["height"], **kw)


def _case(name, H, seed, width=640, height=480, **kw):
    cam_p = {}

    def render(cam, hV, hF, oV, oF, T):
        cam_p.update(cam)
        return O.render_depth(_oparams(cam), hV, hF, oV, oF, T)[0]
    return synth.make_render_case(name, H=H, seed=seed, width=width, height=height, render=render, **kw)


@pytest.mark.parametrize("name,seed", [("ellipse", 3), ("cuboid", 4), ("tless", 5)])
def test_render_depth_bit_exact(ctx, name, seed):
    case = _case(name, 6, seed)
    scene = ctx.render_scene(_params(ctx, case["cam"]), case["depth_m"], case["hand_V"], case["hand_F"])
    op = _oparams(case["cam"])
    for T in (case["gt"], case["poses"][0], case["poses"][3]):
        d, m = ctx.render_depth(scene, case["obj_V"], case["obj_F"], T)
        od, om = O.render_depth(op, case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], T)
        assert np.array_equal(m, om) and np.array_equal(d, od)
    assert om.sum() > 500 and (od < 1.99).sum() > om.sum()       # the object and the hand are both in view
    scene.free()


@pytest.mark.parametrize("name,seed,H", [("ellipse", 3, 96), ("cuboid", 4, 64), ("cylinder", 6, 64)])
def test_reject_by_render_bit_exact(ctx, name, seed, H):
    case = _case(name, H, seed)
    case["poses"][5, :3, 3] = [5.0, 0.0, 0.5]                    # out of view: 0 / 0 = NaN, never kept
    case["poses"][6, :3, 3] = [0.0, 0.0, -0.5]                   # behind the camera
    scene = ctx.render_scene(_params(ctx, case["cam"]), case["depth_m"], case["hand_V"], case["hand_F"])
    wr, order = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], case["poses"])
    owr, oorder = O.reject_by_render(_oparams(case["cam"]), case["depth_m"], case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], case["poses"])
    assert np.array_equal(np.isnan(wr), np.isnan(owr)) and np.isnan(wr[5]) and np.isnan(wr[6])
    assert np.array_equal(wr[~np.isnan(wr)], owr[~np.isnan(owr)])
    assert np.array_equal(order, oorder) and len(order) == max(int(0.3 * H), 10)
    scene.free()


def test_render_without_hand_small_image_and_edges(ctx):
    case = _case("ellipse", 16, 9, width=160, height=120)
    p = _params(ctx, case["cam"], roi_weight=3.0, keep_ratio=0.5)
    op = _oparams(case["cam"], roi_weight=3.0, keep_ratio=0.5)
    scene = ctx.render_scene(p, case["depth_m"])                # no hand meshes
    wr, order = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], case["poses"])
    owr, oorder = O.reject_by_render(op, case["depth_m"], None, None, case["obj_V"], case["obj_F"], case["poses"])
    assert np.array_equal(wr, owr, equal_nan=True) and np.array_equal(order, oorder) and len(order) == 10   # max(int(0.5 * 16), 10)
    # an object that fills the whole image, one that straddles the border
    T = case["gt"].copy(); T[:3, 3] = [0.0, 0.0, 0.12]
    T2 = case["gt"].copy(); T2[:3, 3] = [-0.12, 0.08, 0.35]
    for pose in (T, T2):
        d, m = ctx.render_depth(scene, case["obj_V"], case["obj_F"], pose)
        od, om = O.render_depth(op, None, None, case["obj_V"], case["obj_F"], pose)
        assert np.array_equal(d, od) and np.array_equal(m, om)
    w2, _ = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], np.stack([T, T2]))
    ow2, _ = O.reject_by_render(op, case["depth_m"], None, None, case["obj_V"], case["obj_F"], np.stack([T, T2]))
    assert np.array_equal(w2, ow2, equal_nan=True)
    wr0, order0 = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], case["poses"][:0])
    assert len(wr0) == 0 and len(order0) == 0
    scene.free()
    with pytest.raises(Exception):
        ctx.render_scene(ctx.render_params(width=0), case["depth_m"])


def test_reject_by_render_full_size_properties(ctx):
    """1024 hypotheses at 640 x 480: permuting the hypotheses permutes the wrong ratios bit for bit, a sample matches the oracle,
    and the true pose ranks among the best explanations of the image"""
    case = _case("ellipse", 1024, 12, mesh_level=3)
    case["poses"][0] = case["gt"]
    scene = ctx.render_scene(_params(ctx, case["cam"]), case["depth_m"], case["hand_V"], case["hand_F"])
    wr, order = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], case["poses"])
    perm = np.random.default_rng(0).permutation(1024)
    wr2, _ = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], case["poses"][perm])
    assert np.array_equal(wr2, wr[perm], equal_nan=True)
    sample = np.arange(0, 1024, 64)
    owr, _ = O.reject_by_render(_oparams(case["cam"]), case["depth_m"], case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], case["poses"][sample])
    assert np.array_equal(wr[sample], owr, equal_nan=True)
    assert len(order) == 307 and (wr < wr[0]).sum() < 16 and 0 in order
    scene.free()


def test_pose_estimator_reject_by_render_method(ctx):
    """the host mirror: PoseEstimator.rejectByRender keeps max(0.3 N, 10) hypotheses in ascending wrong ratio"""
    from hop_b200.pose_estimator import PoseEstimator
    case = _case("cuboid", 48, 21)
    pe = PoseEstimator(ctx, {})
    pe.registerMesh(case["obj_V"], case["obj_F"], "object")
    pe.setPoseHypos(case["poses"])
    cam = case["cam"]
    K = [[cam["fx"], 0, cam["cx"]], [0, cam["fy"], cam["cy"]], [0, 0, 1]]
    hand = dict(component_status={"finger": True, "palm": False}, meshes={"finger": (case["hand_V"], case["hand_F"]), "palm": (case["hand_V"], case["hand_F"])},
                tf_in_base={"finger": np.eye(4), "palm": np.eye(4)}, handbase_in_cam=np.eye(4))
    pe.rejectByRender(0.3, hand, case["depth_m"], K, {"render_roi_weight": 2.0, "render_keep_hypo": 0.3})
    owr, oorder = O.reject_by_render(_oparams(cam), case["depth_m"], case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], case["poses"])
    assert [h._id for h in pe._pose_hypos] == list(oorder) and len(oorder) == 14
    assert np.array_equal(np.array([h._wrong_ratio for h in pe._pose_hypos], np.float32), owr[oorder])


def test_wide_image_uses_fewer_hypotheses_per_cta(ctx):
    """1280 x 720 with hypotheses whose tiles span the whole width: the walk kernel's shared-memory rows no longer fit 32 hypotheses
    per CTA; results stay bit-exact"""
    case = _case("ellipse", 40, 14, width=1280, height=720)
    near = case["gt"].copy(); near[:3, 3] = [0.0, 0.0, 0.13]                  # fills the image
    poses = np.concatenate([case["poses"][:36], np.stack([near, near, case["gt"], case["gt"]])])
    scene = ctx.render_scene(_params(ctx, case["cam"]), case["depth_m"], case["hand_V"], case["hand_F"])
    wr, order = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], poses)
    owr, oorder = O.reject_by_render(_oparams(case["cam"]), case["depth_m"], case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], poses)
    assert np.array_equal(wr, owr, equal_nan=True) and np.array_equal(order, oorder)
    d, m = ctx.render_depth(scene, case["obj_V"], case["obj_F"], near)
    cols = np.nonzero(m.any(0))[0]
    assert cols[-1] - cols[0] > 700                                           # the tile is wide enough that fewer than 32 fit (stride > 750)
    scene.free()
