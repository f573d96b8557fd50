"""Shared by the ICP parity tests: which hypotheses have a reference answer that can be matched at all?

The reference's ICP (Utils::runICP -> PCL's float LM with a forward-difference Jacobian) is, on some inputs, not a function of
its input at the 1 mm / 1 deg level: a perturbation of the hypothesis by 1e-7 m, or a different rounding inside the LM (the
reference tree's own Eigen LM versus its C restatement, which agree step for step on well-posed problems), sends the trajectory
to a different stopping point.  No implementation -- including a second build of the reference -- can be "within 1 mm / 1 deg
of the reference" there.  The tests therefore assert the bound on every hypothesis whose reference answer is reproducible, and
list the others:

  unstable : the oracle's own result moves by more than 0.25 mm / 0.25 deg under four 1e-7 m perturbations of the hypothesis
             translation or when its LM step is the reference tree's Eigen LM (oracle/_ref) instead of the C restatement;
  weak     : the first iteration's point-to-plane Jacobian in the reference's parametrisation (translation + rotation about the
             camera origin, columns scaled to unit norm as MINPACK does) has a singular-value ratio below 5e-3: along the weak direction the reference's LM stops where the rounding noise of its forward-difference
             Jacobian (h = sqrt(eps)|x_j|: as small as 1e-8) stalls it, not where the objective does.
"""
import numpy as np

from hop_b200 import synth
from oracle import cpu_oracle as O

SPREAD_T, SPREAD_R = 2.5e-4, 0.25
WEAK_COND = 5e-3


def reference_spread(s, sn, m, mn, hyp, ref, name=None, max_iter=10, n_pert=4, eps=1e-7, seed=0, **icp_kw):
    rng = np.random.default_rng(seed)
    err = synth.pose_error if name is None else (lambda a, b: synth.pose_error_sym(a, b, name))
    wt, wr = np.zeros(len(hyp)), np.zeros(len(hyp))
    runs = []
    for _ in range(n_pert):
        h2 = hyp.copy()
        h2[:, :3, 3] += rng.normal(0, eps, (len(hyp), 3)).astype(np.float32)
        runs.append(O.refine_by_icp(s, sn, m, mn, h2, max_iter=max_iter, **icp_kw)[0])
    if O.ref() is not None:
        O.use_ref_lm(True)
        try:
            runs.append(O.refine_by_icp(s, sn, m, mn, hyp, max_iter=max_iter, **icp_kw)[0])
        finally:
            O.use_ref_lm(False)
    for r in runs:
        dt, dr = err(r, ref)
        wt, wr = np.maximum(wt, dt), np.maximum(wr, dr)
    return wt, wr


def first_iteration_conditioning(s, sn, m, mn, hyp, dist=0.01, angle=45.0):
    from scipy.spatial import cKDTree
    out = np.zeros(len(hyp))
    for i, pose in enumerate(hyp):
        mx, mnn = O.transform_cloud(pose, m, mn)
        d, j = cKDTree(mx).query(s)
        keep = (d.astype(np.float32) ** 2 <= np.float32(dist) ** 2) & (np.sum(sn * mnn[j], 1) > np.cos(np.radians(angle)))
        if keep.sum() < 6:
            continue
        p, n = s[keep].astype(np.float64), mnn[j[keep]].astype(np.float64)
        J = np.hstack([n, 2 * np.cross(p, n)])   # the reference's own parametrisation: rotations about the camera origin
        nrm = np.linalg.norm(J, axis=0)
        if nrm.min() <= 0:
            continue
        sv = np.linalg.svd(J / nrm, compute_uv=False)
        out[i] = sv[-1] / sv[0]
    return out


def icp_classes(s, sn, m, mn, hyp, ref, name=None, max_iter=10, **icp_kw):
    """(unstable, weak) boolean masks over the hypotheses"""
    wt, wr = reference_spread(s, sn, m, mn, hyp, ref, name=name, max_iter=max_iter, **icp_kw)
    unstable = ~(wt <= SPREAD_T) | ~(wr <= SPREAD_R)   # (a NaN pose from a perturbed reference run counts as unstable)
    weak = first_iteration_conditioning(s, sn, m, mn, hyp) < WEAK_COND
    return unstable, weak


def assert_icp_bound(got, ref, s, sn, m, mn, hyp, name=None, max_iter=10, pos_tol=1e-3, rot_tol=1.0, sanity=0.9, flags=None,
                     **icp_kw):
    """every hypothesis with a reproducible reference answer is within pos_tol / rot_tol; returns (ok, unstable, weak).
    `sanity` only guards against a vacuous pass (everything exempted); the parity statement is the first assertion."""
    err = synth.pose_error if name is None else (lambda a, b: synth.pose_error_sym(a, b, name))
    dt, dr = err(got, ref)
    ok = (dt <= pos_tol) & (dr <= rot_tol)
    if ok.all():
        unstable = weak = np.zeros(len(ok), bool)
    else:
        unstable, weak = icp_classes(s, sn, m, mn, hyp, ref, name=name, max_iter=max_iter, **icp_kw)
    exempt = unstable | weak
    bad = np.nonzero(~ok & ~exempt)[0]
    assert len(bad) == 0, ("outside 1 mm / 1 deg with a reproducible reference answer", bad, dt[bad], dr[bad])
    assert ok.mean() >= sanity, ("most hypotheses miss the bound: the exemptions would make this test vacuous", ok.mean(), exempt.mean())
    if flags is not None:   # (iterations, converged) of both sides
        it, cv, rit, rcv = flags
        assert np.array_equal(cv[~exempt], rcv[~exempt]), np.nonzero((cv != rcv) & ~exempt)[0]
    return ok, unstable, weak
