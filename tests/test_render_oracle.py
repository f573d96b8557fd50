"""CPU tests of the render-rejection oracle (oracle/hop_oracle_render.c).  PARITY UNPINNED against the reference's OpenGL renderer
(no GL context here): the camera model and the per-pixel rules are checked against closed forms -- a fronto-parallel
quad covers exactly the pixels whose centres fall inside its projection and reads back its depth in rounded millimetres, the
depth_sim convention sx = fx X/Z + cx, sy = fy Y/Z + (height - cy), flips included -- and the comparison / selection rules
against a literal numpy restatement of PoseEstimator.cpp:399-458."""
import numpy as np

from hop_b200 import synth
from oracle import cpu_oracle as O

QUAD_F = np.array([[0, 1, 2], [0, 2, 3]], np.int32)


def _quad(x0, x1, y0, y1, z):
    return np.array([[x0, y0, z], [x1, y0, z], [x1, y1, z], [x0, y1, z]], np.float32)


def test_camera_convention_and_depth_quantisation():
    p = O.render_params()
    V = _quad(-0.05, 0.07, -0.03, 0.04, 0.4321)
    depth, mask = O.render_depth(p, None, None, V, QUAD_F, np.eye(4))
    xs, ys = np.meshgrid(np.arange(p.width) + 0.5, np.arange(p.height) + 0.5)
    sx = lambda X: p.fx * X / 0.4321 + p.cx
    sy = lambda Y: p.fy * Y / 0.4321 + (p.height - p.cy)
    inside = (xs > sx(-0.05) + 0.01) & (xs < sx(0.07) - 0.01) & (ys > sy(-0.03) + 0.01) & (ys < sy(0.04) - 0.01)
    outside = (xs < sx(-0.05) - 0.01) | (xs > sx(0.07) + 0.01) | (ys < sy(-0.03) - 0.01) | (ys > sy(0.04) + 0.01)
    assert np.all(mask[inside] == 1) and np.all(mask[outside] == 0)
    assert np.all(depth[mask == 1] == np.float32(432) / np.float32(1000)) and np.all(depth[mask == 0] == 2.0)   # round(432.1 mm); cleared depth = z_far
    assert abs(int(mask.sum()) - (sx(0.07) - sx(-0.05)) * (sy(0.04) - sy(-0.03))) < 2 * (mask.sum() ** 0.5)


def test_tilted_plane_depth_is_perspective_correct():
    p = O.render_params()
    V = np.array([[-0.1, -0.08, 0.30], [0.1, -0.08, 0.50], [0.1, 0.08, 0.50], [-0.1, 0.08, 0.30]], np.float32)   # z = 0.4 + x
    depth, mask = O.render_depth(p, None, None, V, QUAD_F, np.eye(4))
    ys, xs = np.nonzero(mask)
    rx = (xs + 0.5 - p.cx) / p.fx                       # X/Z of the pixel ray
    z = 0.4 / (1 - rx)                                  # ray meets z = 0.4 + x
    assert np.abs(depth[ys, xs] - np.round(1000 * z) / 1000).max() <= 0.001 + 1e-6   # at most the last rounded millimetre
    assert (np.abs(depth[ys, xs] - np.round(1000 * z) / 1000) > 1e-6).mean() < 0.01


def test_nearest_surface_wins_and_object_needs_strictly_less():
    p = O.render_params()
    hand = _quad(-0.05, 0.05, -0.05, 0.05, 0.30)
    obj = _quad(-0.02, 0.08, -0.02, 0.02, 0.35)
    depth, mask = O.render_depth(p, hand, QUAD_F, obj, QUAD_F, np.eye(4))
    assert mask.sum() > 0 and np.all(depth[mask == 1] == np.float32(0.35)) and np.float32(0.3) in depth
    cx, cy = int(p.cx), int(p.height - p.cy)
    assert mask[cy, cx] == 0 and depth[cy, cx] == np.float32(0.3)            # behind the hand
    depth2, mask2 = O.render_depth(p, hand, QUAD_F, _quad(-0.02, 0.08, -0.02, 0.02, 0.30), QUAD_F, np.eye(4))
    assert mask2[cy, cx] == 0                                                # equal depth: the hand, drawn first, keeps the pixel (GL_LESS)
    # clipped by the near / far planes, winding does not matter, a moved mesh equals a moved pose
    d3, m3 = O.render_depth(p, None, None, _quad(-0.02, 0.02, -0.02, 0.02, 0.05), QUAD_F, np.eye(4))
    assert m3.sum() == 0 and np.all(d3 == 2.0)
    d4, m4 = O.render_depth(p, None, None, obj, QUAD_F[:, ::-1].copy(), np.eye(4))
    T = np.eye(4); T[:3, 3] = [0.01, -0.02, 0.05]
    d5, m5 = O.render_depth(p, None, None, obj, QUAD_F, T)
    d6, m6 = O.render_depth(p, None, None, (obj + T[:3, 3]).astype(np.float32), QUAD_F, np.eye(4))
    assert np.array_equal(m4, depth * 0 + (O.render_depth(p, None, None, obj, QUAD_F, np.eye(4))[1])) and np.array_equal(m5, m6) and np.array_equal(d5, d6)


def _numpy_wrong_ratio(depth_sim, mask, real, roi_weight):
    """PoseEstimator.cpp:399-443, literally (float32 running sums in row-major order)"""
    roi = bg = np.float32(0)
    roi_cnt = bg_cnt = 0
    for sim, ob, r in zip(depth_sim.ravel(), mask.ravel(), real.ravel()):
        if float(r) <= 0.1 or float(r) >= 2.0:
            d = np.float32(2.0)
        elif float(sim) <= 0.1 or float(sim) >= 2.0:
            d = np.float32(2.0)
        else:
            d = np.float32(abs(np.float32(sim - r)))
        if ob:
            roi = np.float32(roi + d); roi_cnt += 1
        else:
            bg = np.float32(bg + d); bg_cnt += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float32(np.float32(np.float32(roi_weight) * roi) / np.float32(roi_cnt)) + np.float32(bg / np.float32(bg_cnt))


def test_wrong_ratio_and_selection_follow_the_reference_loop():
    p = O.render_params(width=160, height=120, fx=154.149, fy=154.149, cx=76.907, cy=59.922)
    case = synth.make_render_case("ellipse", H=24, seed=3, width=160, height=120,
                                  render=lambda cam, hV, hF, oV, oF, T: O.render_depth(p, hV, hF, oV, oF, T)[0])
    case["poses"][5, :3, 3] = [5.0, 0, 0.5]                     # out of view: owns no pixel -> 0/0 = NaN, sorted last
    wr, order = O.reject_by_render(p, case["depth_m"], case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], case["poses"])
    for h in (0, 7, 5):
        d, m = O.render_depth(p, case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], case["poses"][h])
        want = _numpy_wrong_ratio(d, m, case["depth_m"], p.roi_weight)
        assert (np.isnan(want) and np.isnan(wr[h])) or want == wr[h]
    assert np.isnan(wr[5]) and 5 not in order
    assert len(order) == max(int(p.keep_ratio * 24), 10)
    assert np.all(np.diff(wr[order]) >= 0)                      # ascending, the queue's pop order
    wr_gt, _ = O.reject_by_render(p, case["depth_m"], case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], case["gt"][None])
    assert wr_gt[0] <= np.nanmin(wr) + 1e-3                     # the true pose explains the image best
