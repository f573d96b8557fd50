"""CPU experiment: how close does the moment-replayed LM (csrc/lm_replay.cuh, compiled for the host) get to the reference's
float LM (a) on single solves and (b) over whole ICP trajectories (oracle ICP with the LM step swapped)?"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
from hop_b200 import synth
from oracle import cpu_oracle as O

L = C.CDLL(os.path.join(ROOT, "tests", "support", "liblmr_host.so"))


def icp_with(backend, s, sn, m, mn, hyp, max_iter=10):
    lib = O.lib()
    if backend == "replay":
        lib.hop_oracle_set_lm_backend(C.cast(L.hop_lmr_point_to_plane, C.c_void_p))
    elif backend == "eigen":
        O.use_ref_lm(True)
    else:
        lib.hop_oracle_set_lm_backend(None)
    try:
        return O.refine_by_icp(s, sn, m, mn, hyp, max_iter=max_iter)
    finally:
        lib.hop_oracle_set_lm_backend(None)


if __name__ == "__main__":
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    L.hop_lmr_set_acc_mode(mode)
    tot = bad = 0
    for name, ns, nm, seed, kw in [("ellipse", 600, 3000, 51, dict(rot_sigma_deg=3.0, trans_sigma=0.003)),
                                   ("cuboid", 800, 5000, 51, dict(rot_sigma_deg=3.0, trans_sigma=0.003)),
                                   ("cylinder", 500, 2000, 51, dict(rot_sigma_deg=3.0, trans_sigma=0.003)),
                                   ("tless", 700, 4000, 51, dict(rot_sigma_deg=3.0, trans_sigma=0.003)),
                                   ("ellipse", 2000, 10000, 7, dict()),
                                   ("cuboid", 2000, 10000, 8, dict()),
                                   ("tless", 2000, 10000, 9, dict())]:
        m, mn = synth.make_model(name, nm, seed=1)
        s, sn, conf, gt = synth.make_scene(name, ns, seed=seed)
        hyp = synth.make_hypotheses(gt, 256, seed=seed + 1, random_frac=0.0, **kw)
        t0 = time.time()
        ref, rit, rcv = icp_with("c", s, sn, m, mn, hyp)
        t1 = time.time()
        got, it, cv = icp_with("replay", s, sn, m, mn, hyp)
        t2 = time.time()
        dt, dr = synth.pose_error_sym(got, ref, name)
        ok = (dt <= 1e-3) & (dr <= 1.0)
        tot += len(ok); bad += int((~ok).sum())
        print(f"{name:9s} ns={ns} nm={nm}: ok {ok.mean():.4f}  max dt {dt.max()*1e3:.4f} mm  max dr {dr.max():.4f} deg  "
              f"p99 dt {np.percentile(dt,99)*1e3:.4f} dr {np.percentile(dr,99):.4f}  iters equal {np.mean(it==rit):.3f} conv equal {np.mean(cv==rcv):.3f}"
              f"  (oracle {t1-t0:.1f}s replay {t2-t1:.1f}s)", flush=True)
    print("total", tot, "outside", bad)
