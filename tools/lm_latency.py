"""Latency of K4's inner solver alone (hop_debug_lm_solve, one warp per problem, <= one problem per SM): cycles per solve against
the number of function evaluations the LM took, on the first-iteration moment matrices of bench-like hypotheses.  GPU box only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hop_b200
from hop_b200 import synth
import test_gpu_lm as T

sup = os.path.join(ROOT, "tests", "support")
subprocess.run(["make", "-C", sup], check=True, capture_output=True)
L = C.CDLL(os.path.join(sup, "liblmr_host.so"))
L.hop_lmr_moments.argtypes = [T._f32p, T._f32p, T._f32p, C.c_int, T._f64p]
L.hop_lmr_stats.argtypes = [C.c_void_p, C.c_int]
L.hop_lmr_solve_moments.restype = C.c_int
L.hop_lmr_solve_moments.argtypes = [T._f64p, T._f32p, C.POINTER(C.c_int)]
ctx = hop_b200.Context(0)
for name, kw in [("ellipse", dict()), ("cuboid", dict(rot_sigma_deg=15.0, trans_sigma=0.015)), ("tless", dict(rot_sigma_deg=5.0, trans_sigma=0.005))]:
    mats = T._moment_sets(L, name, 2000, 10000, 7, 140, kw)[:140]
    sums = np.stack([T._pack(A) for A in mats])
    ctx.debug_lm_solve(sums, with_cycles=True)
    x, nfev, st, cyc = ctx.debug_lm_solve(sums, with_cycles=True)
    # host work counters per problem
    cold = []
    out = (C.c_longlong * 5)()
    for A in mats:
        Af = np.ascontiguousarray(A.astype(np.float32).astype(np.float64))
        L.hop_lmr_stats(out, 1)
        xx = np.zeros(6, np.float32); nf = C.c_int(0)
        L.hop_lmr_solve_moments(Af.reshape(-1), xx, C.byref(nf))
        L.hop_lmr_stats(out, 1)
        cold.append((out[1], out[2], out[3], out[4]))
    cold = np.array(cold)
    ok = st >= 0
    outer = np.maximum(cold[ok, 0], 1)
    print(f"{name}: n {ok.sum()}  nfev mean {nfev[ok].mean():.1f} max {nfev[ok].max()}  cycles mean {cyc[ok].mean():.0f} max {cyc[ok].max()}  "
          f"cycles/nfev {cyc[ok].sum() / nfev[ok].sum():.0f}  cycles/outer(host count) {cyc[ok].sum() / outer.sum():.0f}")
    # least squares: cycles ~ a*outer + b*trials + c*cold_calls + d*qrsolv
    Xm = np.column_stack([cold[ok, 0], cold[ok, 1], cold[ok, 2], cold[ok, 3], np.ones(ok.sum())]).astype(np.float64)
    coef, *_ = np.linalg.lstsq(Xm, cyc[ok].astype(np.float64), rcond=None)
    print("   fit cycles = %.0f*outer + %.0f*trials + %.0f*lmpar_cold + %.0f*qrsolv + %.0f" % tuple(coef))
