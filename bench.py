#!/usr/bin/env python
"""bench.py -- pose hypotheses/s through the hot path (ICP refinement K4 + LCP scoring K5 + winner selection).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|headline|C3|tiny] [--impl ours|reference]

One "step" = one depth frame: the frame's scene cloud and its batch of H pose hypotheses go through
  [scene NN-grid build] -> icp_refine (K4) -> lcp_score (K5) -> top-K winners (-> one all-gather of winners when N>1).
`value`  : hypotheses/s with the scene cloud and hypotheses already resident in HBM (CUDA events, L2 flushed between steps).
`e2e`    : the same through the host-buffer C ABI (hop_cloud_update + hop_icp_refine + hop_lcp_score: what
           PoseEstimator::refineByICP()/selectBest() call), pinned host inputs, H2D + D2H inside the timed region.
`roofline`: dominant kernel (icp_fused_kernel): algorithmic bytes / its own CUDA-event time vs the measured HBM peak.
`cpu_baseline`: the oracle port of the reference algorithm on this box's host cores (bounded sample), rank 0, N=1.
Multi-GPU: frames are independent -> every rank processes its own frames (weak scaling), one NCCL all-gather of the
per-rank winner records per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))

METRIC = "pose hypotheses/sec (ICP-refined + LCP-scored)"
UNIT = "hypotheses/s"
TOPK = 16
# dram__bytes_read.sum + dram__bytes_write.sum of one icp_fused_kernel launch (a whole C2 batch), from the ncu --set full
# captures summarised in profiles/r01_ncu_icp_fused_kernel_C2.txt and _headline.txt; other workloads were not captured -> null
TRAFFIC_BYTES_PER_LAUNCH = {"C2": 30083840, "headline": 36006144}
N_FRAMES = 4  # distinct synthetic frames cycled through the steps (a step = `frames_per_step` consecutive frames, default 1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--pipeline", type=int, default=0, help="0 = persistent fused kernel (default), 1 = per-iteration moments + solve kernels")
    ap.add_argument("--solver", type=int, default=0, help="0 = replay of the reference's LM (parity path), 1 = Gauss-Newton, 2 = exact per-iteration minimiser")
    ap.add_argument("--cpu-sample", type=int, default=0, help="hypotheses in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_frames(wl, rank):
    from hop_b200 import synth
    m, mn = synth.make_model(wl["model"], wl["n_model"], seed=1)
    frames = []
    for f in range(N_FRAMES):
        seed = 1000 * 2 + 17 * rank + f
        s, sn, conf, gt = synth.make_scene(wl["model"], wl["n_scene"], seed=seed)
        hy = synth.make_hypotheses(gt, wl["H"], seed=seed + 500)
        frames.append(dict(xyz=s, nrm=sn, conf=conf, gt=gt, hyp=hy))
    return (m, mn), frames


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows)}


def host_threads():
    """all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: not what the CPU arm should get)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_run(model, frame, wl, sample, threads=0):
    """reference algorithm (oracle port: kd-tree rebuilt per hypothesis, LM point-to-plane ICP, LCP) on host cores"""
    from oracle import cpu_oracle as O
    threads = threads or host_threads()
    m, mn = model
    hy = frame["hyp"][:sample]
    t0 = time.perf_counter()
    ref, it, cv = O.refine_by_icp(frame["xyz"], frame["nrm"], m, mn, hy, max_iter=wl["max_iter"], nthreads=threads)
    best, sc = O.select_best(frame["xyz"], frame["nrm"], m, mn, ref, nthreads=threads)
    dt = time.perf_counter() - t0
    return len(hy) / dt, dt, threads


def run_reference(args, wl, rank, world, out):
    """--impl reference: the reference's own CPU algorithm for the path (PCL is not installable -> oracle port)."""
    if rank != 0:
        return
    from oracle import cpu_oracle as O
    model, frames = make_frames(wl, 0)
    # size the per-step sample so the whole run stays within a few minutes
    probe = min(8, wl["H"])
    rate, _, threads = cpu_baseline_run(model, frames[0], wl, probe)
    budget_s = 120.0 / max(args.steps + args.warmup, 1)
    sample = int(max(probe, min(wl["H"], rate * budget_s)))
    for w in range(args.warmup):
        cpu_baseline_run(model, frames[w % N_FRAMES], wl, sample)
    t_total = 0.0
    for k in range(args.steps):
        r, dt, threads = cpu_baseline_run(model, frames[k % N_FRAMES], wl, sample)
        t_total += dt
    value = sample * args.steps / t_total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, **wl, "sample_hypotheses_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} of {wl['H']} hypotheses per step, all {threads} host threads (OpenMP over hypotheses)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, ...) write to file descriptor 1 too:
    keep a private copy of the real stdout for the result line and point fd 1 at stderr for everybody else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse()
    out = claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from hop_b200 import synth
    wl = synth.workload(args.workload)

    if args.impl == "reference":
        run_reference(args, wl, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    import hop_b200

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    ctx = hop_b200.Context(local_rank)  # raises when the CUDA library / device is missing: no fallback
    # a real (non-default) stream: everything libhop enqueues and every timing event lives on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    model_np, frames = make_frames(wl, rank)
    H, ns, nm = wl["H"], wl["n_scene"], wl["n_model"]
    FPS = int(wl.get("frames_per_step", 1))   # frames of one rank in one step (C4: 16), one all-gather of winners per step
    model = ctx.upload_cloud(*model_np)
    icp_p = ctx.icp_params(max_iter=wl["max_iter"], solver=args.solver, pipeline=args.pipeline)
    lcp_p = ctx.lcp_params()
    # per-model structures are built once (like weights): not part of a frame's step
    g_icp = model.prepare_nn(icp_p.max_dist)
    g_lcp = model.prepare_nn(lcp_p.dist)

    # ---- device-resident inputs for `value` ----
    scenes = [ctx.upload_cloud(f["xyz"], f["nrm"], f["conf"]) for f in frames]
    d_hyp = [torch.from_numpy(hop_b200.capi.poses_to_colmajor(f["hyp"])).to(dev) for f in frames]
    d_pose = torch.empty((H, 16), dtype=torch.float32, device=dev)
    d_iters = torch.zeros(H, dtype=torch.int32, device=dev)
    d_conv = torch.zeros(H, dtype=torch.int32, device=dev)
    d_score = torch.zeros(H, dtype=torch.float32, device=dev)
    d_send = torch.zeros(FPS * TOPK * 80, dtype=torch.uint8, device=dev)
    d_recv = torch.zeros(world * FPS * TOPK * 80, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_value(k, evs=None, on_frame=None):
        if evs: evs[0].record(stream)
        for f in range(FPS):
            fi = (k * FPS + f) % N_FRAMES
            d_pose.copy_(d_hyp[fi])  # ICP refines in place; keep the inputs pristine (device-to-device, 64 KB per 1024)
            sc = scenes[fi]
            sc.drop_nn()  # the scene's reciprocal-NN grid is per-frame work: rebuilt inside the timed step ...
            sc.prepare_lcp_scene(lcp_p)  # ... on the context's second stream, while the ICP runs (the LCP kernel waits for it)
            ctx.icp_refine_dev(sc, model, d_pose.data_ptr(), H, icp_p, d_iters.data_ptr(), d_conv.data_ptr())
            if evs and f == 0: evs[1].record(stream)
            ctx.lcp_score_dev(sc, model, d_pose.data_ptr(), H, lcp_p, d_score.data_ptr())
            if evs and f == 0: evs[2].record(stream)
            ctx.select_topk_dev(d_pose.data_ptr(), d_score.data_ptr(), H, TOPK, d_send.data_ptr() + f * TOPK * 80, id_offset=0, frame=rank * FPS + f)
            if on_frame: on_frame()
        if world > 1:
            dist.all_gather_into_tensor(d_recv, d_send)
        if evs: evs[3].record(stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    # (every distinct frame is visited at least once: a frame's first visit allocates its scene grid)
    for w in range(max(args.warmup, N_FRAMES)):
        step_value(w)
    barrier()

    # ---- timed: K steps, device time per step from events, L2 flushed before each ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    l0 = ctx.launch_count()
    barrier()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        step_value(k, ev[k])
    barrier()
    launches = ctx.launch_count() - l0
    t_step = np.array([e[0].elapsed_time(e[3]) for e in ev])
    t_icp = np.array([e[0].elapsed_time(e[1]) for e in ev])
    t_lcp = np.array([e[1].elapsed_time(e[2]) for e in ev])
    total_ms = float(t_step.sum())
    iters_mean = float(d_iters.float().mean().item())

    # ---- per-kernel device time: the same K steps again with libhop's event profiling on (CUDA events recorded on
    #      the launching stream around every launch of each kernel family; kept out of the `value` timing) ----
    ctx.profile_enable(True)
    iters_box = [0]

    def count_iters():
        torch.cuda.synchronize()
        iters_box[0] += int(torch.clamp(d_iters, min=1).sum().item())

    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        step_value(k, None, count_iters)
    iters_sum = iters_box[0]
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # ---- e2e through the host-buffer ABI ----
    e2e_scene = ctx.upload_cloud(frames[0]["xyz"], frames[0]["nrm"], frames[0]["conf"])
    pin = []
    for f in frames:
        p = {k: ctx.pinned_array(f[k].shape, np.float32) for k in ("xyz", "nrm", "conf")}
        for k in p:
            p[k][...] = f[k]
        p["hyp"] = ctx.pinned_array((H, 16), np.float32)
        p["hyp"][...] = hop_b200.capi.poses_to_colmajor(f["hyp"])
        p["work"] = ctx.pinned_array((H, 16), np.float32)
        p["iters"] = ctx.pinned_array((H,), np.int32)
        p["conv"] = ctx.pinned_array((H,), np.int32)
        p["scores"] = ctx.pinned_array((H,), np.float32)
        p["win"] = ctx.pinned_array((1,), hop_b200.capi.POSE_REC_DTYPE)
        pin.append(p)
    import ctypes as C
    vp = C.c_void_p

    def ptr(a):
        return a.ctypes.data_as(vp)

    def step_e2e(k):
        best = []
        for f in range(FPS):
            p = pin[(k * FPS + f) % N_FRAMES]
            ctx._check(ctx.L.hop_cloud_update(ctx.h, e2e_scene.handle, ptr(p["xyz"]), ptr(p["nrm"]), ptr(p["conf"]), ns))
            ctx._check(ctx.L.hop_refine_score_select(ctx.h, e2e_scene.handle, model.handle, None, ptr(p["hyp"]), H, C.byref(icp_p), C.byref(lcp_p), 0, 1,
                                                     ptr(p["work"]), ptr(p["scores"]), ptr(p["iters"]), ptr(p["conv"]), ptr(p["win"])))
            best.append(int(p["win"]["id"][0]))  # selectBest's arg-max
        return best

    for w in range(max(args.warmup, N_FRAMES)):
        step_e2e(w)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ms = 0.0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        torch.cuda.synchronize()
        e0.record(stream)
        step_e2e(k)
        e1.record(stream)
        torch.cuda.synchronize()
        e2e_ms += e0.elapsed_time(e1)
    barrier()
    clocks = sampler.summary()
    h2d = FPS * (ns * 7 * 4 + H * 64)               # the frame's cloud + the hypotheses, once
    d2h = FPS * (H * 64 + H * 8 + H * 4 + 80)       # refined poses, iterations + flags, scores, the winner record

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0].item()), float(t[1].item())

    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        value = world * FPS * H * args.steps / (total_ms * 1e-3)
        e2e_value = world * FPS * H * args.steps / (e2e_ms * 1e-3)
        # Dominant kernel: icp_fused_kernel (the whole ICP of the batch, one launch per step).  ALGORITHMIC bytes (SURVEY
        # 8d): every executed ICP iteration of a hypothesis streams the scene and the model once as 2 x float4 = 32 B/pt.
        fused = args.pipeline != 1
        corr_ms, corr_n = prof["icp_fused" if fused else "icp_correspond"]
        corr_bytes = float(iters_sum) * 32.0 * (ns + nm)
        achieved = corr_bytes / (corr_ms * 1e-3) / 1e9
        icp_ms = float(np.mean(t_icp))
        kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items() if v[1]}
        stage_note = "first frame of the step" if FPS > 1 else "the step's frame"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, **wl, "frames_per_rank_per_step": FPS, "topk": TOPK, "l2": "flushed between steps (256 MiB write)",
                       "icp_solver": {0: "reference-lm-replay", 1: "gauss-newton", 2: "exact"}[args.solver],
                       "icp_pipeline": {0: "persistent fused", 1: "moments+solve per iteration"}[args.pipeline], "mean_icp_iterations": iters_mean,
                       "nn_grid_icp": g_icp, "nn_grid_lcp": g_lcp,
                       "stage_ms": {"icp_refine": icp_ms, "lcp_score": float(np.mean(t_lcp)), "step": total_ms / args.steps, "of": stage_note},
                       "kernel_ms": kernels},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "icp_fused_kernel" if fused else "icp_moments_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH.get(args.workload) if fused else None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "algorithmic_bytes_per_launch": corr_bytes / max(corr_n, 1), "kernel_ms_per_launch": corr_ms / max(corr_n, 1),
                         "launches": int(corr_n), "kernel_ms_per_step": corr_ms / args.steps,
                         "note": "algorithmic = 32 B x (N_scene + N_model) per executed ICP iteration per hypothesis (the brute-force "
                                 "streaming model of SURVEY 8d); the kernel itself gathers from an L2-resident voxel grid, so real DRAM "
                                 "traffic (`traffic`, ncu) is far below it; what bounds it is the L1TEX pipe (81.7 % of peak at the headline "
                                 "size, profiles/r01_ncu_icp_fused_kernel_headline.txt)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload: whole frames' batches (or the first hypotheses of one) for >= ~10 s of CPU work
            probe_rate, _, thr = cpu_baseline_run(model_np, frames[0], wl, min(8, H))
            sample = args.cpu_sample or int(max(8, min(H, probe_rate * 12.0)))
            done, spent, k = 0, 0.0, 0
            while spent < 10.0 and k < 64:
                rate, dt, thr = cpu_baseline_run(model_np, frames[k % N_FRAMES], wl, sample)
                done, spent, k = done + sample, spent + dt, k + 1
            line["cpu_baseline"] = {"value": done / spent, "unit": UNIT, "cores": thr, "kind": "port",
                                    "sample": f"{k} x first {sample} of {H} hypotheses (frames cycled), {thr} OpenMP threads, {spent:.1f} s"}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
