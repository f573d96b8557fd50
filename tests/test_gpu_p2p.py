"""mode 1 of hop_icp_refine: Utils::runICP(segment, model, T, max_corres_dist) (Utils.cpp:135-164) = PCL's default point-to-point
ICP with reciprocal correspondences and the SVD (Umeyama) transformation estimation, against its oracle restatement."""
import numpy as np
import pytest

from hop_b200 import synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ns,nm", [("ellipse", 600, 3000), ("cuboid", 800, 5000), ("tless", 700, 4000)])
def test_point_to_point_icp_matches_oracle(ctx, name, ns, nm):
    m, mn = synth.make_model(name, nm, seed=1)
    s, sn, conf, gt = synth.make_scene(name, ns, seed=5)
    hyp = synth.make_hypotheses(gt, 48, seed=6, random_frac=0.0, rot_sigma_deg=3.0, trans_sigma=0.003)
    scene, model = ctx.upload_cloud(s, sn, conf), ctx.upload_cloud(m, mn)
    # (1) the trajectory: a fixed number of iterations (the MSE stop disabled) -- same reciprocal correspondences, same Kabsch steps
    for iters in (1, 5, 20):
        got, it, cv = ctx.icp_refine(scene, model, hyp, ctx.icp_params(mode=1, max_iter=iters, abs_mse_eps=0.0, max_dist=0.01))
        ref, rit, rcv = O.refine_by_icp_p2p(s, m, hyp, max_iter=iters, dist=0.01, abs_mse_eps=0.0)
        dt, dr = synth.pose_error(got, ref)
        assert np.array_equal(cv, rcv) and np.array_equal(it, rit)
        # (one correspondence at a float tie may differ: the scene is queried in the model frame here, the model in the scene frame there)
        assert np.median(dt) < 1e-6 and dt.max() < 5e-5 and np.median(dr) < 0.01 and dr.max() < 0.5, (iters, dt.max(), dr.max())
    # (2) PCL's defaults: up to 100 iterations, stop when |dMSE| < 1e-12 m^2.  Near convergence the MSE of float coordinates moves by
    # about that much per iteration through rounding alone (the reference re-transforms its cloud in float every iteration), so WHERE a
    # run stops along the slow final creep is decided by noise -- in the reference as much as here (a numpy emulation of this kernel's
    # arithmetic reproduces its stop to the iteration).  Same flags, same fixed point for most, all inside the creep band.
    got, it, cv = ctx.icp_refine(scene, model, hyp, ctx.icp_params(mode=1, max_iter=100, abs_mse_eps=1e-12, max_dist=0.01))
    ref, rit, rcv = O.refine_by_icp_p2p(s, m, hyp, max_iter=100, dist=0.01, abs_mse_eps=1e-12)
    assert np.array_equal(cv, rcv)
    dt, dr = synth.pose_error_sym(got, ref, name)
    assert np.median(dt) < 1e-6 and np.median(dr) < 1e-3
    assert np.mean((dt <= 1e-4) & (dr <= 0.05)) >= 0.85 and np.all((dt <= 1.5e-3) & (dr <= 2.0)), (dt.max(), dr.max())
    scene.free(); model.free()


def test_point_to_point_semantics(ctx):
    m, mn = synth.make_model("ellipse", 2000, seed=1)
    s, sn, conf, gt = synth.make_scene("ellipse", 400, seed=7)
    hyp = synth.make_hypotheses(gt, 8, seed=8, random_frac=0.0, rot_sigma_deg=2.0, trans_sigma=0.002)
    scene, model = ctx.upload_cloud(s, sn, conf), ctx.upload_cloud(m, mn)
    # fewer than 3 reciprocal correspondences: not converged, pose unchanged (Utils.cpp:156-163)
    far = hyp.copy(); far[:, :3, 3] += 1.0
    got, it, cv = ctx.icp_refine(scene, model, far, ctx.icp_params(mode=1, max_iter=100, abs_mse_eps=1e-12))
    assert np.all(cv == 0) and np.all(it == 0) and np.allclose(got, far, atol=1e-6)
    # one iteration = one Kabsch step on the reciprocal correspondences of the hypothesis
    got, it, cv = ctx.icp_refine(scene, model, hyp, ctx.icp_params(mode=1, max_iter=1, abs_mse_eps=1e-12))
    ref, rit, rcv = O.refine_by_icp_p2p(s, m, hyp, max_iter=1)
    dt, dr = synth.pose_error(got, ref)
    assert np.all(it == 1) and np.all(cv == 1) and dt.max() < 2e-5 and dr.max() < 0.01
    scene.free(); model.free()
