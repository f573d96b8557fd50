"""K4's inner solver on the device (csrc/lm_replay_warp.cuh: the LM replay laid out over the lanes of a warp) against the scalar
program it restates (csrc/lm_replay.cuh compiled for the host, tests/support -- which tests/test_lm_replay.py pins against the
reference tree's own Eigen LM), on the same float moment matrices."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hop_b200 import synth
from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def lmr():
    sup = os.path.join(HERE, "support")
    subprocess.run(["make", "-C", sup], check=True, capture_output=True)
    L = C.CDLL(os.path.join(sup, "liblmr_host.so"))
    L.hop_lmr_moments.argtypes = [_f32p, _f32p, _f32p, C.c_int, _f64p]
    L.hop_lmr_solve_moments.restype = C.c_int
    L.hop_lmr_solve_moments.argtypes = [_f64p, _f32p, C.POINTER(C.c_int)]
    return L


def _moment_sets(lmr, name, ns, nm, seed, n_hyp, pose_kw):
    """Moment matrices the ICP of `n_hyp` hypotheses meets in its first iteration (float sums, like the kernel's)."""
    from scipy.spatial import cKDTree
    m, mn = synth.make_model(name, nm, seed=1)
    s, sn, conf, gt = synth.make_scene(name, ns, seed=seed)
    hyp = synth.make_hypotheses(gt, n_hyp, seed=seed + 1, random_frac=0.0, **pose_kw)
    out = []
    for pose in hyp:
        mx, mnn = O.transform_cloud(pose, m, mn)
        d, j = cKDTree(mx).query(s)
        keep = (d ** 2 <= 0.01 ** 2) & (np.sum(sn * mnn[j], 1) > np.cos(np.radians(45)))
        if keep.sum() < 6:
            continue
        A = np.zeros(169, np.float64)
        lmr.hop_lmr_moments(np.ascontiguousarray(s[keep]), np.ascontiguousarray(mx[j[keep]]), np.ascontiguousarray(mnn[j[keep]]), int(keep.sum()), A)
        out.append(A.reshape(13, 13))
    return out


def _pack(A):
    sums = np.zeros(96, np.float32)
    k = 0
    for i in range(13):
        for j in range(i, 13):
            sums[k] = np.float32(A[i, j]); k += 1
    return sums


@pytest.mark.parametrize("name,kw", [("ellipse", dict(rot_sigma_deg=3.0, trans_sigma=0.003)), ("cuboid", dict(rot_sigma_deg=3.0, trans_sigma=0.003)),
                                     ("tless", dict(rot_sigma_deg=5.0, trans_sigma=0.005)), ("cuboid", dict(rot_sigma_deg=15.0, trans_sigma=0.015)),
                                     ("cylinder", dict(rot_sigma_deg=5.0, trans_sigma=0.005))])
def test_warp_lm_follows_the_scalar_program(ctx, lmr, name, kw):
    mats = _moment_sets(lmr, name, 800, 4000, 11, 64, kw)
    assert len(mats) >= 32
    x_dev, nfev_dev, st_dev = ctx.debug_lm_solve(np.stack([_pack(A) for A in mats]))
    n_same_nfev = 0
    worst = 0.0
    for k, A in enumerate(mats):
        Af = np.ascontiguousarray(A.astype(np.float32).astype(np.float64))   # the float sums the kernel sees
        x = np.zeros(6, np.float32)
        nfev = C.c_int(0)
        st = lmr.hop_lmr_solve_moments(Af.reshape(-1), x, C.byref(nfev))
        if st == -1:
            assert st_dev[k] == -1
            continue
        assert st_dev[k] in (1, 2, 3, 4, 5, 6, 7, 8) and np.isfinite(x_dev[k]).all()
        n_same_nfev += int(nfev.value == nfev_dev[k])
        # same objective at the stopping point (the stopping point itself moves along weak directions with the last bit of a norm)
        y0, y1 = O.warp_y13(x), O.warp_y13(x_dev[k])
        f0, f1 = float(y0 @ Af @ y0), float(y1 @ Af @ y1)
        f_init = float(Af[12, 12])   # the objective at x = 0
        cut_off = max(nfev.value, int(nfev_dev[k])) >= 400   # maxfev ended a run that was still creeping: wherever it happened to be
        assert abs(f1 - f0) <= (0.1 if cut_off else 2e-3) * abs(f0) + 1e-5 * f_init, (k, f0, f1, f_init, nfev.value, nfev_dev[k])
        worst = max(worst, float(np.abs(x - x_dev[k]).max()))
    # device division / square root are the approximate SFU ones (<= 2 ulp): most runs still take the very same steps
    assert n_same_nfev >= 0.5 * len(mats), (n_same_nfev, len(mats))
    if name == "cuboid" and kw["rot_sigma_deg"] < 5:
        assert worst < 2e-4, worst


def test_warp_lm_rank_deficient_jacobian(ctx, lmr):
    """Source points all at the origin: the three rotation columns of the Jacobian are exactly zero, the Cholesky factor has zero pivots
    and lmpar takes MINPACK's rank-deficient branch -- on the device the one place that falls back to the scalar lmpar_iterate, fed from
    the factor in shared memory.  Same translation as the host program; the rotation parameters stay 0."""
    rng = np.random.default_rng(5)
    sums, mats = [], []
    for k in range(16):
        n = rng.normal(size=(200, 3)).astype(np.float32)
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        src = np.zeros((200, 3), np.float32)
        tgt = (rng.normal(scale=0.004, size=3).astype(np.float32) + rng.normal(scale=1e-4, size=(200, 3)).astype(np.float32))
        A = np.zeros(169, np.float64)
        lmr.hop_lmr_moments(src, np.ascontiguousarray(tgt), np.ascontiguousarray(n), 200, A)
        mats.append(A.reshape(13, 13)); sums.append(_pack(mats[-1]))
    x_dev, nfev_dev, st_dev = ctx.debug_lm_solve(np.stack(sums))
    for k, A in enumerate(mats):
        Af = np.ascontiguousarray(A.astype(np.float32).astype(np.float64))
        x = np.zeros(6, np.float32)
        nfev = C.c_int(0)
        st = lmr.hop_lmr_solve_moments(Af.reshape(-1), x, C.byref(nfev))
        assert st >= 1 and st_dev[k] >= 1 and np.isfinite(x_dev[k]).all()
        assert np.all(x_dev[k][3:] == 0) and np.all(x[3:] == 0)
        assert np.abs(x_dev[k][:3] - x[:3]).max() < 2e-6, (k, x, x_dev[k])
