"""Time hop_s4pcs_plan_create_gpu (device PPF membership + the serial RNG replay) at a few scene sizes.  GPU box only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
import hop_b200
from hop_b200 import capi, synth

ctx = hop_b200.Context(0)
m, mn = synth.make_model("ellipse", 10000, seed=1)
sub = slice(None, None, max(1, len(m) // 400))
keys = ctx.ppf_table(m[sub], mn[sub])
for ns in (959, 2000, 10000, 50000):
    s, sn, conf, gt = synth.make_scene("ellipse", ns, seed=2)
    opt = capi.s4pcs_options(sample_size=100)
    capi.S4pcsPlan(s, sn, conf, m, mn, keys, opt, ctx=ctx)
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        pl = capi.S4pcsPlan(s, sn, conf, m, mn, keys, opt, ctx=ctx)
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"scene {ns:6d} points: plan {dt:8.2f} ms   (base_ok {int(np.sum(pl.get()['base_ok']))} of 30 trials)", flush=True)
