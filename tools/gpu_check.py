"""Quick GPU-side parity report (development aid; the real gates are tests/ -m gpu)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
import numpy as np
import hop_b200
from hop_b200 import synth
from oracle import cpu_oracle as O

def main():
    ctx = hop_b200.Context(0)
    name, ns, nm, H = "ellipse", 500, 2000, 64
    if len(sys.argv) > 1: ns, nm, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    m, mn = synth.make_model(name, nm, 1)
    s, sn, conf, gt = synth.make_scene(name, ns, 2)
    hy = synth.make_hypotheses(gt, H, 3)
    scene = ctx.upload_cloud(s, sn, conf); model = ctx.upload_cloud(m, mn)
    # NN grid exactness
    print("grid icp:", model.prepare_nn(0.01)); print("grid lcp:", model.prepare_nn(0.001))
    rng = np.random.default_rng(0)
    q = (m[rng.integers(0, nm, 4000)] + rng.normal(0, 0.004, (4000, 3))).astype(np.float32)
    for R in (0.01, 0.001):
        gi, gd = model.nn_query(R, q)
        oi, od = O.nn(m, q, use_kdtree=False)
        oi = np.where(od <= R * R, oi, -1)
        bad = (gi != oi)
        # ties / boundary
        print(f"nn R={R}: mismatches {bad.sum()} / {len(q)}; found {np.sum(gi>=0)}; max d2 err {np.abs(np.where(gi>=0, gd-od, 0)).max():.3e}")
    for solver in (0, 1):
        t = time.time()
        ref_p, ref_it, ref_cv = O.refine_by_icp(s, sn, m, mn, hy, nthreads=0)
        t_cpu = time.time() - t
        t = time.time()
        p = ctx.icp_params(solver=solver)
        out, it, cv = ctx.icp_refine(scene, model, hy, p)
        t_gpu = time.time() - t
        dt, dr = synth.pose_error(out, ref_p)
        print(f"ICP solver={solver}: cpu {t_cpu:.3f}s gpu {t_gpu:.4f}s  iters equal {np.mean(it==ref_it):.3f} conv equal {np.mean(cv==ref_cv):.3f}")
        print(f"   dt mm: max {dt.max()*1e3:.4f} p99 {np.percentile(dt,99)*1e3:.4f} median {np.median(dt)*1e3:.5f};  drot deg: max {dr.max():.4f} p99 {np.percentile(dr,99):.4f} median {np.median(dr):.5f}")
        w = np.argsort(-dt)[:5]
        print("   worst:", [(int(i), round(float(dt[i]*1e3),3), round(float(dr[i]),3), int(it[i]), int(ref_it[i]), int(cv[i]), int(ref_cv[i])) for i in w])
    best, sc_ref = O.select_best(s, sn, m, mn, ref_p, nthreads=0)
    sc = ctx.lcp_score(scene, model, ref_p)
    rel = np.abs(sc - sc_ref) / np.maximum(np.abs(sc_ref), 1e-3)
    print(f"LCP: best ref {best} gpu {int(np.argmax(sc))}; max rel err {rel.max():.3e}; max abs {np.abs(sc-sc_ref).max():.3e}; ref range {sc_ref.min():.2f}..{sc_ref.max():.2f}")
    top = ctx.select_topk(ref_p, sc, 4)
    print("topk ids", top["id"], "scores", top["score"])
    print("launches", ctx.launch_count())

if __name__ == "__main__":
    main()
