#!/usr/bin/env python
"""Stage benches next to bench.py's headline line: K1 (hand-state overlap search) and Super4PCS (K2a pair extraction,
K2b congruent-set search, K3 verification) at the BASELINE.json sizes, one JSON line per stage.

  python tools/bench_stages.py [--steps K] [--warmup W] [--sizes C2|C5]

Per stage: units/s end to end through the host-buffer C ABI (what Hand::matchOneComponentPSO / PoseEstimator::runSuper4pcs
call; H2D/D2H inside), the kernel's own CUDA-event time (libhop's per-kernel profiling on the launching stream) and the
roofline entry from SURVEY 8(d)'s algorithmic bytes:  K1  32 (N_f + N_h) + 32 N_w + 8  per hand state,
K3  16 (nQ + N_s) + 84  per congruent quadrilateral.  No oracle, no reference: the PPF key set comes from hop_compute_ppf.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))

SIZES = {
    # S hand states, finger / hand-scene points; Super4PCS: scene points, model points handed to the matcher, sample size
    "C2": dict(S=4096, n_finger=300, n_hand=3000, n_scene=2000, n_model=10000, sample=100),
    "C5": dict(S=16384, n_finger=400, n_hand=5000, n_scene=50000, n_model=50000, sample=400),
}


def ppf_keys(xyz, nrm, stride=1):
    """unique PPF keys of all point pairs (what computePPF.cpp:86-107 tabulates), through the library's own hop_compute_ppf"""
    import ctypes as C
    import hop_b200
    L = hop_b200.load_library()
    xyz = np.ascontiguousarray(xyz[::stride], np.float32)
    nrm = np.ascontiguousarray(nrm[::stride], np.float32)
    n = len(xyz)
    key = (C.c_int32 * 4)()
    fp = C.POINTER(C.c_float)
    seen = set()
    px = [xyz[i].ctypes.data_as(fp) for i in range(n)]
    pn = [nrm[i].ctypes.data_as(fp) for i in range(n)]
    f = L.hop_compute_ppf
    for i in range(n):
        for j in range(i + 1, n):
            f(px[i], pn[i], px[j], pn[j], key)
            seen.add((key[0], key[1], key[2], key[3]))
    return np.array(sorted(seen), np.int32).reshape(-1, 4)


def collision_stage(ctx, args, out, timed, peak):
    """physics pruning: PoseEstimator::rejectByCollisionOrNonTouching for a batch of hypotheses (hop_reject_by_collision)"""
    from hop_b200 import synth
    H = 1024 if args.sizes == "C2" else 16384
    n_model = 600 if args.sizes == "C2" else 2000           # _model is the 5 mm cloud (a few hundred points in the reference)
    case = synth.make_collision_case("ellipse", H=H, seed=51, n_model=n_model, mesh_level=2 if args.sizes == "C2" else 3)
    obj = ctx.upload_mesh(case["obj_V"], case["obj_F"])
    fm = [ctx.upload_mesh(v, f) for v, f in zip(case["finger_V"], case["finger_F"])]
    fc = [ctx.upload_cloud(p) for p in case["finger_pts"]]
    scene, hand, model = ctx.upload_cloud(case["scene_xyz"]), ctx.upload_cloud(case["hand_xyz"]), ctx.upload_cloud(case["model_xyz"])
    params = ctx.collision_params(case["params"])
    res = {}

    def run():
        res["r"] = ctx.reject_by_collision(obj, fm, fc, scene, hand, model, case["poses"], params)

    dt, prof = timed(run)
    k_ms = prof["sdf"][0] / max(prof["sdf"][1], 1)
    keep, reason, diag = res["r"]
    # point-triangle tests the reference's brute-force equivalent would do for the steps each hypothesis reached
    nfo, nff = len(case["obj_F"]), sum(len(f) for f in case["finger_F"])
    npf = sum(len(p) for p in case["finger_pts"])
    tests = float(np.sum((reason != 1) * 1.0) * 0 + len(reason) * nfo + np.sum(reason != 1) * nfo + np.sum(~np.isin(reason, [1, 2])) * npf * nfo
                  + np.sum(np.isin(reason, [0, 5, 6])) * len(case["model_xyz"]) * nff)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import cpu_oracle as O
        thr = max(1, len(os.sched_getaffinity(0)))
        n_s = min(H, 1024)
        sub = dict(case)
        sub["poses"] = case["poses"][:n_s]
        t0 = time.perf_counter()
        okeep, oreason, _ = O.reject_by_collision(sub)
        tc = time.perf_counter() - t0
        cpu = {"value": n_s / tc, "unit": "hypotheses/s", "cores": thr, "kind": "port", "same_decisions": float(np.mean(oreason == reason[:n_s])),
               "sample": f"first {n_s} of {H} hypotheses, {tc:.2f} s (restated igl signed distance, brute force over faces, OpenMP over hypotheses; "
                         "the reference rebuilds an AABB tree per hypothesis instead)"}
    print(json.dumps({"stage": "rejectByCollisionOrNonTouching (hop_reject_by_collision: signed distances + the whole decision, one launch)",
                      "metric": "hypotheses pruned/sec", "value": H / (k_ms * 1e-3), "unit": "hypotheses/s", "cpu_baseline": cpu,
                      "e2e": {"value": H / dt, "unit": "hypotheses/s", "ms_per_call": dt * 1e3, "h2d_bytes": 64 * H, "d2h_bytes": 48 * H},
                      "config": {"sizes": args.sizes, "H": H, "object_faces": nfo, "finger_faces": nff, "finger_points": npf, "n_model": len(case["model_xyz"]),
                                 "n_scene": len(case["scene_xyz"]), "n_hand": len(case["hand_xyz"]), "kept": int(keep.sum()),
                                 "reasons": np.bincount(reason, minlength=7).tolist()},
                      "kernel_ms": k_ms, "dtype": "f32", "point_triangle_tests_per_launch": tests,
                      "gtests_per_s": tests / (k_ms * 1e-3) / 1e9,
                      "note": "ALU-bound (meshes and clouds stay in L1/L2): the figure of merit is point-triangle tests per second, not HBM bytes"}),
          file=out, flush=True)


def hand_removal_stage(ctx, args, out, timed):
    """hand-point removal with confidences (HandT42::removeSurroundingPointsAndAssignProbability) on the device"""
    from hop_b200 import synth
    n = 6000 if args.sizes == "C2" else 50000
    case = synth.make_hand_removal_case(seed=5, n_scene=n, n_link=400 if args.sizes == "C2" else 2000)
    p = ctx.hand_removal_params(case["handbase_in_cam"], case["finger_1_2_in_handbase"], case["finger_2_2_in_handbase"], case["min_z"], case["near_hand_dist"])
    scene = ctx.upload_cloud(case["scene_xyz"], case["scene_nrm"])
    lc = [ctx.upload_cloud(l) for l in case["links"]]
    res = {}

    def run():
        o = ctx.remove_hand_points(scene, lc, case["kinds"], p)
        res["n"] = o.n
        o.free()

    dt, prof = timed(run)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import cpu_oracle as O
        t0 = time.perf_counter()
        ox, _, _ = O.remove_hand_points(case["scene_xyz"], case["scene_nrm"], case["links"], case["kinds"], p)
        tc = time.perf_counter() - t0
        cpu = {"value": n / tc, "unit": "scene points/s", "cores": max(1, len(os.sched_getaffinity(0))), "kind": "port", "ms_per_call": tc * 1e3,
               "same_count": bool(len(ox) == res["n"]), "sample": "the whole cloud (brute-force exact nearest neighbour per link, OpenMP over points)"}
    print(json.dumps({"stage": "hand-point removal + confidences (hop_remove_hand_points)", "metric": "scene points/sec", "value": n / dt, "unit": "scene points/s",
                      "cpu_baseline": cpu, "e2e": {"value": n / dt, "unit": "scene points/s", "ms_per_call": dt * 1e3},
                      "config": {"sizes": args.sizes, "n_scene": n, "links": len(lc), "points_per_link": len(case["links"][0]), "kept": res["n"]},
                      "kernel_ms": None, "dtype": "f32"}), file=out, flush=True)


def render_stage(ctx, args, out, timed):
    """render-based rejection: PoseEstimator::rejectByRender for a batch of hypotheses (hop_reject_by_render)"""
    from hop_b200 import synth
    H = 1024 if args.sizes == "C2" else 16384
    cam = {}

    def render(c, hV, hF, oV, oF, T):   # the "real" depth image of the synthetic frame: the product's own renderer, true pose
        cam.update(c)
        sc0 = ctx.render_scene(ctx.render_params(**c), np.zeros((c["height"], c["width"]), np.float32), hV, hF)
        d = ctx.render_depth(sc0, oV, oF, T)[0]
        sc0.free()
        return d
    case = synth.make_render_case("ellipse", H=H, seed=12, mesh_level=3, render=render)
    p = ctx.render_params(**cam)
    res = {}
    t0 = time.perf_counter()
    scene = ctx.render_scene(p, case["depth_m"], case["hand_V"], case["hand_F"])
    t_scene = time.perf_counter() - t0

    def run():
        res["r"] = ctx.reject_by_render(scene, case["obj_V"], case["obj_F"], case["poses"])

    dt, prof = timed(run)
    k_ms = prof["render"][0] / max(prof["render"][1], 1)
    wr, order = res["r"]
    cpu = None
    if not args.no_cpu_baseline:   # the only use of the oracle here: the CPU baseline leg
        from oracle import cpu_oracle as O
        thr = max(1, len(os.sched_getaffinity(0)))
        n_s = min(H, 256)
        t0 = time.perf_counter()
        owr, _ = O.reject_by_render(O.render_params(**cam), case["depth_m"], case["hand_V"], case["hand_F"], case["obj_V"], case["obj_F"], case["poses"][:n_s])
        tc = time.perf_counter() - t0
        cpu = {"value": n_s / tc, "unit": "hypotheses/s", "cores": thr, "kind": "port", "bit_identical": bool(np.array_equal(owr, wr[:n_s], equal_nan=True)),
               "sample": f"first {n_s} of {H} hypotheses, {tc:.2f} s (software rasteriser + the reference's comparison loop, OpenMP over hypotheses; "
                         "the reference renders serially through one OpenGL context)"}
    npx = cam["width"] * cam["height"]
    print(json.dumps({"stage": "rejectByRender (hop_reject_by_render: bbox + rasterise + the reference's row-major float sums, 4 launches)",
                      "metric": "hypotheses rendered and compared/sec", "value": H / (k_ms * 1e-3), "unit": "hypotheses/s", "cpu_baseline": cpu,
                      "e2e": {"value": H / dt, "unit": "hypotheses/s", "ms_per_call": dt * 1e3, "h2d_bytes": 64 * H, "d2h_bytes": 12 * H},
                      "config": {"sizes": args.sizes, "H": H, "image": [cam["width"], cam["height"]], "object_faces": len(case["obj_F"]),
                                 "hand_faces": len(case["hand_F"]), "kept": int(len(order)), "scene_setup_ms": t_scene * 1e3},
                      "kernel_ms": k_ms, "dtype": "f32 sums, f64 depth interpolation, int64 coverage",
                      "roofline": {"bound": "hbm", "kernel": "walk_kernel", "unit": "GB/s", "achieved": H * 8.0 * npx / (k_ms * 1e-3) / 1e9, "peak": None,
                                   "note": "algorithmic = the reference's loop reads 2 floats per pixel per hypothesis; the sums are sequential by definition "
                                           "(latency-bound: one dependent FADD per pixel per hypothesis)"}}), file=out, flush=True)
    scene.free()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="all", help="all | collision | render | physics (collision + render)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sizes", default="C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")   # the JSON lines; fd 1 itself goes to stderr (the compiled reference matcher prints its progress)
    os.dup2(2, 1)
    import hop_b200
    from hop_b200 import capi, hand, synth
    sz = SIZES[args.sizes]
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ctx = hop_b200.Context(0)

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        ctx.sync()
        ctx.profile_enable(True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        ctx.sync()
        dt = (time.perf_counter() - t0) / args.steps
        prof = ctx.profile_read()
        ctx.profile_enable(False)
        return dt, prof

    if args.stages == "hand_removal":
        hand_removal_stage(ctx, args, out, timed)
        ctx.close()
        return
    if args.stages in ("collision", "render", "physics"):
        if args.stages != "render":
            collision_stage(ctx, args, out, timed, peak)
        if args.stages != "collision":
            render_stage(ctx, args, out, timed)
        ctx.close()
        return

    # ---------------- K1: S joint angles of one finger link ----------------
    case = synth.make_hand_case(seed=5, n_finger=sz["n_finger"], n_hand=sz["n_hand"])
    prop = hand.FingerProperty(case["finger_xyz"], case["scalars"]["num_division"])
    p = hand.finger_params(prop, case["scalars"])
    finger = ctx.upload_cloud(case["finger_xyz"], case["finger_nrm"])
    scene = ctx.upload_cloud(case["scene_xyz"], case["scene_nrm"])
    lookup = ctx.upload_cloud(case["lookup_xyz"], case["lookup_nrm"])
    nosw = ctx.upload_cloud(case["noswivel_xyz"], case["noswivel_nrm"])
    thetas = np.deg2rad(np.linspace(0, 120, sz["S"]))
    res = {}

    def k1():
        res["k1"] = ctx.hand_overlap(finger, scene, nosw, p, thetas, lookup)

    dt, prof = timed(k1)
    S, nf, nh, nw = sz["S"], len(case["finger_xyz"]), len(case["scene_xyz"]), len(case["noswivel_xyz"])
    k_ms = prof["hand_overlap"][0] / max(prof["hand_overlap"][1], 1)
    bytes_state = 32 * (nf + nh) + 32 * nw + 8
    ach = S * bytes_state / (k_ms * 1e-3) / 1e9
    cpu1 = None
    if not args.no_cpu_baseline:   # the objFuncPSO restatement (oracle port), OpenMP over states, on a bounded sample of the same grid
        from oracle import cpu_oracle as O
        thr = max(1, len(os.sched_getaffinity(0)))
        n_s = min(S, 512)
        t0 = time.perf_counter()
        O.hand_overlap(p, case["finger_xyz"], case["finger_nrm"], case["scene_xyz"], case["lookup_nrm"], case["noswivel_xyz"], thetas[:: S // n_s][:n_s], nthreads=thr)
        tc = time.perf_counter() - t0
        cpu1 = {"value": n_s / tc, "unit": "states/s", "cores": thr, "kind": "port", "sample": f"{n_s} of {S} states, {tc:.2f} s (kd-tree built once per call; the reference deep-copies it per thread per generation)"}
    print(json.dumps({"stage": "K1 hand_overlap (objFuncPSO over a dense grid of joint angles)", "metric": "hand states/sec", "value": S / (k_ms * 1e-3), "cpu_baseline": cpu1,
                      "unit": "states/s", "e2e": {"value": S / dt, "unit": "states/s", "ms_per_call": dt * 1e3, "h2d_bytes": 8 * S, "d2h_bytes": 8 * S + 4},
                      "config": {"sizes": args.sizes, "S": S, "n_finger": nf, "n_scene_hand": nh, "n_noswivel": nw,
                                 "best_theta_deg": float(np.rad2deg(thetas[res["k1"][1]])), "true_theta_deg": float(np.rad2deg(case["theta_true"]))},
                      "kernel_ms": k_ms, "dtype": "f32 (f64 cost)",
                      "roofline": {"bound": "hbm", "kernel": "hand_overlap_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                   "traffic": None, "algorithmic_bytes_per_launch": S * bytes_state}}), file=out, flush=True)

    # ---------------- Super4PCS: plan on the host (untimed: it replays the reference's RNG), K2a + K2b + K3 on the device ----------------
    m, mn = synth.make_model("ellipse", sz["n_model"], seed=1)
    s, sn, conf, gt = synth.make_scene("ellipse", sz["n_scene"], seed=2)
    t0 = time.perf_counter()
    keys = ppf_keys(m, mn, stride=max(1, len(m) // 400))          # the 5 mm "ppf_density" model of computePPF.cpp: ~400 points
    t_keys = time.perf_counter() - t0
    t0 = time.perf_counter()
    plan = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options(sample_size=sz["sample"]))
    t_plan = time.perf_counter() - t0

    def s4():
        res["s4"] = ctx.super4pcs_run(plan, capacity=400000)

    plan_k = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options(sample_size=sz["sample"], keep_intermediates=1))
    ctx.super4pcs_run(plan_k, capacity=400000)      # once, untimed, with the intermediates kept: how many quadrilaterals K3 verifies
    info = plan_k.sizes()
    dt, prof = timed(s4)
    n_hyp = len(res["s4"][1])
    steps = args.steps
    v_ms = prof["verify_lcp"][0] / steps
    pr_ms = prof["s4pcs_pairs"][0] / steps
    jn_ms = prof["s4pcs_join"][0] / steps
    M, nQ = int(info["quads"]), int(info["nQ"])
    bytes_quad = 16 * (nQ + len(s)) + 84
    ach = (M * bytes_quad / (v_ms * 1e-3) / 1e9) if v_ms > 0 and M > 0 else 0.0
    adi = float("nan")
    if n_hyp:   # ADI of the best-LCP hypothesis (the reference's own metric, scripts/eval_utils.py:181-200; symmetric objects flip freely)
        from scipy.spatial import cKDTree
        best = res["s4"][0][int(np.argmax(res["s4"][1]))].astype(np.float64)
        sub = m[::5].astype(np.float64)
        adi = float(cKDTree(sub @ gt[:3, :3].T + gt[:3, 3]).query(sub @ best[:3, :3].T + best[:3, 3])[0].mean())
    cpu2 = None
    if not args.no_cpu_baseline and args.sizes == "C2":   # (bounded: the CPU matcher needs minutes at the C5 sizes)
        from oracle import cpu_oracle as O
        if O.ref() is not None and hasattr(O.ref(), "hop_ref_s4pcs_run"):   # the reference's own compiled matcher (oracle/_ref), all host threads
            thr = max(1, len(os.sched_getaffinity(0)))
            best_t, n_ref = 1e30, 0
            for _ in range(3):
                t0 = time.perf_counter()
                r = O.ref_super4pcs(s, sn, conf, m, mn, keys, sample_size=sz["sample"], nthreads=thr)
                best_t = min(best_t, time.perf_counter() - t0)
                n_ref = len(r["lcp"])
            cpu2 = {"value": n_ref / best_t, "unit": "emitted hypotheses/s", "cores": thr, "kind": "reference", "ms_per_call": best_t * 1e3,
                    "sample": f"whole registration (init + {len(r['trials'])} trials + verification), best of 3, {n_ref} hypotheses, {len(r['quads'])} quadrilaterals"}
    print(json.dumps({"stage": "Super4PCS device stages (K2a pairs, K2b congruent sets, K3 verify), all trials of a frame in one call", "cpu_baseline": cpu2,
                      "metric": "congruent quadrilaterals verified/sec", "value": (M / (v_ms * 1e-3)) if v_ms > 0 else 0.0, "unit": "quads/s",
                      "e2e": {"value": n_hyp / dt, "unit": "emitted hypotheses/s", "ms_per_call": dt * 1e3},
                      "config": {"sizes": args.sizes, "n_scene": len(s), "n_model": len(m), "plan": info, "hypotheses_emitted": n_hyp,
                                 "ppf_keys": int(len(keys)), "host_plan_ms": t_plan * 1e3, "host_ppf_table_ms": t_keys * 1e3,
                                 "best_lcp_hypothesis_adi_mm": adi * 1e3},
                      "kernel_ms": {"k2a_pairs": pr_ms, "k2b_join": jn_ms, "k3_verify": v_ms}, "dtype": "f32 / int32",
                      "roofline": {"bound": "hbm", "kernel": "verify_lcp_kernel", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                   "traffic": None, "algorithmic_bytes_per_launch": M * bytes_quad}}), file=out, flush=True)
    # ---------------- the frame's front end: depth image -> object-segment cloud ----------------
    Kc = (616.5961303710938, 616.59619140625, 307.6278076171875, 239.68692016601562)
    dense, _ = synth.make_model("ellipse", 400000, seed=7)
    hb = np.eye(4); hb[:3, 3] = [0.15, 0.0, 0.38]
    gtf = np.eye(4); gtf[:3, :3] = synth._rot_from_rotvec(np.array([0.4, -0.7, 0.3])); gtf[:3, 3] = [0.0, 0.005, 0.35]
    Pd = dense.astype(np.float64) @ gtf[:3, :3].T + gtf[:3, 3]
    uu = np.round(Pd[:, 0] * Kc[0] / Pd[:, 2] + Kc[2]).astype(int); vv = np.round(Pd[:, 1] * Kc[1] / Pd[:, 2] + Kc[3]).astype(int)
    dimg = np.full((480, 640), np.inf); np.minimum.at(dimg, (vv, uu), Pd[:, 2]); dimg[~np.isfinite(dimg)] = 0.45   # object in front of a wall
    depth = np.round(dimg * 1000 + np.random.default_rng(1).normal(0, 0.5, dimg.shape)).astype(np.uint16)
    Tf = np.linalg.inv(hb).astype(np.float32)
    fpar = ctx.frame_params(K=Kc, cam_in_handbase=Tf)
    res["frame"] = ctx.frame_to_scene(depth, fpar)

    def fr():
        res["frame"] = ctx.frame_to_scene(depth, fpar, scene=res["frame"][0])

    dt, prof = timed(fr)
    cpu4 = None
    tool = os.path.join(ROOT, "icra20-hand-object-pose_b200", "host", "host_tool")
    if not args.no_cpu_baseline and os.path.exists(tool):
        import subprocess, tempfile, cv2
        with tempfile.TemporaryDirectory() as td:
            cv2.imwrite(td + "/d.png", depth); np.savetxt(td + "/T.txt", Tf)
            r = subprocess.run([tool, "frame", td + "/d.png", *map(repr, Kc), td + "/T.txt", td + "/o.bin"], capture_output=True, text=True)
            ms = [float(l.split()[1]) for l in r.stdout.splitlines() if l.startswith("frame_ms")]
        if ms:
            cpu4 = {"value": 1e3 / ms[0], "unit": "frames/s", "cores": 1, "kind": "port", "ms_per_call": ms[0],
                    "sample": "one 640x480 frame through host/cloud.cpp frameToObjectSegment (the PCL chain restated, single thread like the reference)"}
    print(json.dumps({"stage": "frame front end on the device (hop_frame_to_scene: back-projection, 2 voxel grids, crop, normals)", "metric": "depth frames/sec",
                      "value": 1.0 / dt, "unit": "frames/s", "cpu_baseline": cpu4,
                      "e2e": {"value": 1.0 / dt, "unit": "frames/s", "ms_per_call": dt * 1e3, "h2d_bytes": int(depth.nbytes), "d2h_bytes": 0},
                      "config": {"image": "640x480 uint16 mm", "stage_counts": [int(x) for x in res["frame"][1]]},
                      "kernel_ms": prof["frame"][0] / max(prof["frame"][1], 1)}), file=out, flush=True)

    # ---------------- the whole frame through the drop-in executable: main_realdata_auto <config.yaml> <repeat> ----------------
    main_bin = os.path.join(ROOT, "icra20-hand-object-pose_b200", "host", "main_realdata_auto")
    if os.path.exists(main_bin) and args.sizes == "C2":
        import subprocess, tempfile, cv2
        from scipy.spatial import cKDTree
        import contextlib
        keep = os.environ.get("HOP_KEEP_FRAME_DIR")   # keep the frame's config + inputs (e.g. to profile the executable on them)
        if keep:
            os.makedirs(keep, exist_ok=True)
        with (contextlib.nullcontext(keep) if keep else tempfile.TemporaryDirectory()) as td:
            cfg = f"""cam_K: [{Kc[0]}, 0.0, {Kc[2]}, 0.0, {Kc[1]}, {Kc[3]}, 0.0, 0.0, 1.0]
cam1_in_leftarm: [0.0,0.0,0.0,0.0,0.0,0.0,1.0]
handbase_in_palm: [1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1]
out_dir: {td}
rgb_path: {td}/rgb.png
depth_path: {td}/depth.png
palm_in_baselink: {td}/palm_in_base.txt
leftarm_in_base: {td}/arm_left.txt
model_name: ellipse
object_model_path: {td}/ellipse.ply
ppf_path: {td}/ppf_ellipse
object_symmetry:
  ellipse:
    x: 180
    y: 180
    z: 180
lcp:
  dist: 0.001
  normal_angle: 10
pose_estimator_high_confidence_thres: 0.8
icp_dist_thres: 0.01
icp_angle_thres: 45
super4pcs_sample_size: 100
super4pcs_overlap: 0.2
super4pcs_delta: 0.003
super4pcs_dispersion: 0.5
super4pcs_success_quadrilaterals: 10
"""
            open(td + "/cfg.yaml", "w").write(cfg)
            np.savetxt(td + "/arm_left.txt", np.eye(4)); np.savetxt(td + "/palm_in_base.txt", hb)
            mm, mmn = synth.make_model("ellipse", 60000, seed=8)
            with open(td + "/ellipse.ply", "w") as f:
                f.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\nend_header\n" % len(mm))
                np.savetxt(f, np.concatenate([mm, mmn], 1), fmt="%.9g")
            dobj = np.full((480, 640), np.inf); np.minimum.at(dobj, (vv, uu), Pd[:, 2]); dobj[~np.isfinite(dobj)] = 0
            dobj_mm = np.round(dobj * 1000).astype(np.uint16)
            cv2.imwrite(td + "/depth.png", dobj_mm)
            res["frame_obj"] = ctx.frame_to_scene(dobj_mm, fpar)[0]
            # the same frame with the physics + render stages on: object_mesh_path readable (no hand model ships: object-only scene)
            mV, mF = synth.make_mesh("ellipse", 3)
            with open(td + "/ellipse.obj", "w") as f:
                f.write("".join("v %.9g %.9g %.9g\n" % tuple(v) for v in mV) + "".join("f %d %d %d\n" % tuple(t + 1) for t in mF))
            open(td + "/cfg_phys.yaml", "w").write(cfg + f"object_mesh_path: {td}/ellipse.obj\npose_estimator_use_physics: true\nrender_roi_weight: 2.0\nrender_keep_hypo: 0.3\n")
            r2 = subprocess.run([main_bin, td + "/cfg_phys.yaml", "6"], capture_output=True, text=True, timeout=600)
            tl2 = [l for l in r2.stdout.splitlines() if l.startswith("timing_ms")]
            if r2.returncode == 0 and tl2:
                w2 = tl2[-1].split()
                tm2 = {w2[i]: float(w2[i + 1]) for i in range(1, len(w2) - 1, 2)}
                est2 = np.loadtxt(td + "/model2scene.txt")
                sub2 = mm[::20].astype(np.float64)
                adi2 = float(cKDTree(sub2 @ gtf[:3, :3].T + gtf[:3, 3]).query(sub2 @ est2[:3, :3].T + est2[:3, 3])[0].mean())
                print(json.dumps({"stage": "one frame through the drop-in executable with rejectByCollisionOrNonTouching + rejectByRender on (object mesh given; 6th pass)",
                                  "metric": "frames/sec, depth image -> best pose", "value": 1e3 / tm2["total"], "unit": "frames/s",
                                  "e2e": {"value": 1e3 / tm2["total"], "unit": "frames/s", "ms_per_call": tm2["total"]},
                                  "config": {"stage_ms": tm2, "adi_mm": adi2 * 1e3}, "kernel_ms": None, "cpu_baseline": None}), file=out, flush=True)
            else:
                print("main_realdata_auto (physics) failed:", r2.returncode, r2.stdout[-800:], r2.stderr[-800:], file=sys.stderr)
            r = subprocess.run([main_bin, td + "/cfg.yaml", "6"], capture_output=True, text=True, timeout=600)
            tl = [l for l in r.stdout.splitlines() if l.startswith("timing_ms")]
            if r.returncode == 0 and tl:
                w_ = tl[-1].split()
                tm = {w_[i]: float(w_[i + 1]) for i in range(1, len(w_) - 1, 2)}
                est = np.loadtxt(td + "/model2scene.txt")
                sub = mm[::20].astype(np.float64)
                adi = float(cKDTree(sub @ gtf[:3, :3].T + gtf[:3, 3]).query(sub @ est[:3, :3].T + est[:3, 3])[0].mean())
                cpu5 = None
                if not args.no_cpu_baseline:
                    # the same frame through the CPU side stage by stage: host front end (cloud.cpp), the reference's own compiled
                    # Super4PCS matcher, host clustering, the ICP / LCP port on the <= 100 clusters (all host threads where the
                    # reference uses OpenMP)
                    from oracle import cpu_oracle as O
                    thr = max(1, len(os.sched_getaffinity(0)))
                    seg_xyz, seg_nrm, seg_conf = res["frame_obj"].download()
                    m5, m5n = synth.make_model("ellipse", 644, seed=8)
                    m1, m1n = synth.make_model("ellipse", 10000, seed=8)
                    stage = {}
                    tool = os.path.join(ROOT, "icra20-hand-object-pose_b200", "host", "host_tool")
                    np.savetxt(td + "/T.txt", Tf)
                    rr = subprocess.run([tool, "frame", td + "/depth.png", *map(repr, Kc), td + "/T.txt", td + "/o.bin"], capture_output=True, text=True)
                    fm = [float(l.split()[1]) for l in rr.stdout.splitlines() if l.startswith("frame_ms")]
                    stage["front_end"] = fm[0] if fm else float("nan")
                    if O.ref() is not None and hasattr(O.ref(), "hop_ref_s4pcs_run"):
                        kk = ppf_keys(m5, m5n)
                        t0 = time.perf_counter(); rs = O.ref_super4pcs(seg_xyz, seg_nrm, seg_conf, m5, m5n, kk, sample_size=100, nthreads=thr); stage["super4pcs"] = (time.perf_counter() - t0) * 1e3
                        t0 = time.perf_counter(); keep = capi.cluster_poses(rs["poses"], rs["lcp"], 30.0, 0.015, (180.0, 180.0, 180.0)); stage["cluster"] = (time.perf_counter() - t0) * 1e3
                        cl = rs["poses"][keep[:100]]
                        t0 = time.perf_counter(); rp, _, _ = O.refine_by_icp(seg_xyz, seg_nrm, m5, m5n, cl, nthreads=thr); stage["icp"] = (time.perf_counter() - t0) * 1e3
                        t0 = time.perf_counter(); O.select_best(seg_xyz, seg_nrm, m1, m1n, rp, nthreads=thr); stage["select"] = (time.perf_counter() - t0) * 1e3
                        tot = sum(stage.values())
                        cpu5 = {"value": 1e3 / tot, "unit": "frames/s", "cores": thr, "kind": "reference (Super4PCS) + port (front end, ICP, LCP)", "ms_per_call": tot,
                                "sample": "the same frame, stage by stage", "stage_ms": stage}
                print(json.dumps({"stage": "one frame through the drop-in executable (main_realdata_auto <cfg> 6: stage times of the 6th pass)", "metric": "frames/sec, depth image -> best pose",
                                  "value": 1e3 / tm["total"], "unit": "frames/s", "e2e": {"value": 1e3 / tm["total"], "unit": "frames/s", "ms_per_call": tm["total"]},
                                  "config": {"stage_ms": tm, "adi_mm": adi * 1e3, "note": "front end, K2a/K2b/K3, both clusterPoses, K4 on <= 100 clusters, K5; the Super4PCS base planner (host, replays the reference's RNG) is inside 'super4pcs'"},
                                  "kernel_ms": None, "cpu_baseline": cpu5}), file=out, flush=True)
            else:
                print("main_realdata_auto failed:", r.returncode, r.stdout[-800:], r.stderr[-800:], file=sys.stderr)

    # ---------------- clusterPoses: host loop vs the device version (identical keep list) ----------------
    n_cl = 20000 if args.sizes == "C2" else 65536
    hyp = synth.make_hypotheses(gt, n_cl, seed=3, rot_sigma_deg=20, trans_sigma=0.02, random_frac=0.3)
    sc = np.random.default_rng(0).random(n_cl).astype(np.float32)

    def cl():
        res["cl"] = ctx.cluster_poses(hyp, sc, 30.0, 0.015)

    dt, prof = timed(cl)
    cpu3 = None
    if not args.no_cpu_baseline and args.sizes == "C2":
        t0 = time.perf_counter()
        ref_keep = capi.cluster_poses(hyp, sc, 30.0, 0.015)
        tc = time.perf_counter() - t0
        cpu3 = {"value": n_cl / tc, "unit": "hypotheses/s", "cores": 1, "kind": "port", "ms_per_call": tc * 1e3,
                "sample": "the whole batch, hop_cluster_poses (the reference's sequential loop, Euler angles hoisted)", "same_keep_list": bool(np.array_equal(ref_keep, res["cl"]))}
    print(json.dumps({"stage": "clusterPoses(30 deg, 15 mm) on the device (hop_cluster_poses_gpu)", "metric": "hypotheses clustered/sec", "value": n_cl / dt,
                      "unit": "hypotheses/s", "cpu_baseline": cpu3, "e2e": {"value": n_cl / dt, "unit": "hypotheses/s", "ms_per_call": dt * 1e3},
                      "config": {"n": n_cl, "clusters": int(len(res["cl"]))}, "kernel_ms": prof["cluster"][0] / max(prof["cluster"][1], 1)}), file=out, flush=True)
    hand_removal_stage(ctx, args, out, timed)
    collision_stage(ctx, args, out, timed, peak)
    render_stage(ctx, args, out, timed)
    ctx.close()


if __name__ == "__main__":
    main()
