"""hop_cloud_prepare_nn_async (-m gpu): the scene grid of hop_lcp_score built ahead on the context's second stream, concurrent
with hop_icp_refine.  The bar is bit-identical results to the implicit (same-stream) build, frame after frame on one cloud handle."""
import numpy as np
import pytest

from hop_b200 import synth

pytestmark = pytest.mark.gpu


def _frame(name, ns, nm, H, seed):
    m, mn = synth.make_model(name, nm, seed=1)
    s, sn, conf, gt = synth.make_scene(name, ns, seed=seed)
    return m, mn, s, sn, conf, synth.make_hypotheses(gt, H, seed=seed + 1, rot_sigma_deg=1.0, trans_sigma=0.001)


def test_prefetched_scene_grid_gives_identical_poses_and_scores(ctx):
    m, mn = synth.make_model("ellipse", 6000, seed=1)
    model = ctx.upload_cloud(m, mn)
    icp_p, lcp_p = ctx.icp_params(max_iter=10), ctx.lcp_params(dist=0.002, angle_deg=15.0)
    plain = ctx.upload_cloud(*_frame("ellipse", 900, 6000, 8, 3)[2:5])
    ahead = ctx.upload_cloud(*_frame("ellipse", 900, 6000, 8, 3)[2:5])
    for k in range(6):   # the same two handles refilled frame after frame, sizes changing (grids rebuilt, buffers regrown)
        _, _, s, sn, conf, hyp = _frame("ellipse", 700 + 300 * (k % 3), 6000, 96, seed=40 + k)
        plain.update(s, sn, conf)
        ref_pose, ref_it, ref_cv = ctx.icp_refine(plain, model, hyp, icp_p)
        ref_score = ctx.lcp_score(plain, model, ref_pose, lcp_p)
        ahead.update(s, sn, conf)
        ahead.prepare_lcp_scene(lcp_p)
        if k == 4:
            ahead.prepare_lcp_scene(lcp_p)   # asking twice is harmless
        got_pose, got_it, got_cv = ctx.icp_refine(ahead, model, hyp, icp_p)
        got_score = ctx.lcp_score(ahead, model, got_pose, lcp_p)
        assert np.array_equal(got_pose, ref_pose) and np.array_equal(got_it, ref_it) and np.array_equal(got_cv, ref_cv)
        assert np.array_equal(got_score, ref_score) and ref_score.max() > 5
    # a prefetch nobody consumes: the next update and the free still order themselves after it
    ahead.prepare_nn_async(0.004)
    ahead.update(s[:100], sn[:100], conf[:100])
    ahead.prepare_nn_async(0.004)
    idx, d2 = ahead.nn_query(0.004, s[:50])
    assert np.array_equal(idx, np.arange(50)) and np.all(d2 == 0)
    ahead.prepare_nn_async(0.006)
    plain.free(); ahead.free(); model.free()


def test_prefetch_errors(ctx):
    import ctypes as C
    assert ctx.L.hop_cloud_prepare_nn_async(None, None, 0.001, 0.0) != 0
    one = ctx.upload_cloud(np.array([[0.1, 0.2, 0.3]], np.float32), np.array([[0, 0, 1]], np.float32))
    assert ctx.L.hop_cloud_prepare_nn_async(ctx.h, one.handle, C.c_float(0.0), C.c_float(0.0)) != 0   # like hop_cloud_prepare_nn: bad radius
    assert b"radius" in ctx.L.hop_last_error(ctx.h)
    # the context still works on its main stream afterwards
    s, sn, conf, gt = synth.make_scene("cuboid", 300, seed=5)
    c = ctx.upload_cloud(s, sn, conf)
    c.prepare_nn_async(0.003)
    idx, d2 = c.nn_query(0.003, s[:20])
    assert np.array_equal(idx, np.arange(20))
    c.free(); one.free()
