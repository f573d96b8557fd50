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
    p = ctx.icp_params(mode=1, max_iter=100, abs_mse_eps=1e-12, max_dist=0.01)
    got, it, cv = ctx.icp_refine(scene, model, hyp, p)
    ref, rit, rcv = O.refine_by_icp_p2p(s, m, hyp, max_iter=100, dist=0.01, abs_mse_eps=1e-12)
    assert np.array_equal(cv, rcv)
    # The Kabsch step has a closed form and the two runs follow the same correspondences bit for bit (median difference 1e-7 m)
    # until a float tie flips one.  The reference's stop, |dMSE| < 1e-12 m^2 over up to 100 iterations, then ends a run wherever
    # the last bits of the MSE settle -- along the weak directions of a partial view that can be tenths of a millimetre apart.
    # As for mode 0 (tests/parity_util.py) the bound is asserted where the reference's own answer is reproducible: its result
    # under three 1e-7 m perturbations of the hypothesis stays within 0.25 mm / 0.25 deg.
    dt, dr = synth.pose_error_sym(got, ref, name)
    assert np.median(dt) < 1e-6 and np.median(dr) < 1e-3
    ok = (dt <= 1e-3) & (dr <= 1.0)
    close = (dt <= 1e-4) & (dr <= 0.05)
    if not close.all():
        rng = np.random.default_rng(0)
        wt, wr = np.zeros(len(hyp)), np.zeros(len(hyp))
        for _ in range(3):
            h2 = hyp.copy(); h2[:, :3, 3] += rng.normal(0, 1e-7, (len(hyp), 3)).astype(np.float32)
            r2, _, _ = O.refine_by_icp_p2p(s, m, h2, max_iter=100, dist=0.01, abs_mse_eps=1e-12)
            a, b = synth.pose_error_sym(r2, ref, name)
            wt, wr = np.maximum(wt, a), np.maximum(wr, b)
        unstable = ~(wt <= 2.5e-4) | ~(wr <= 0.25)
        assert np.all(close | unstable), (np.nonzero(~close & ~unstable)[0], dt.max(), dr.max())
    assert ok.mean() >= 0.9
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
