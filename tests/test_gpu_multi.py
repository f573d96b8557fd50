"""More than one GPU in one box (skipped on a single-GPU box; run under `gpurun --gpus 2`):
  * two hop_ctx on two devices in ONE process (what "one context per host thread / GPU" promises): per-device kernel attributes,
    device guard at every entry point -- identical results on both, with the caller's current device left alone;
  * the path's one collective through the C ABI (hop_comm_init + hop_gather_winners[_dev], NCCL bound at run time): two ranks shard
    a hypothesis batch, gather their winner records and merge them to the single-GPU answer."""
import os
import sys

import numpy as np
import pytest

from hop_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")


@needs2
def test_two_contexts_on_two_devices_in_one_process():
    import torch
    import hop_b200
    m, mn = synth.make_model("ellipse", 3000, seed=1)
    s, sn, conf, gt = synth.make_scene("ellipse", 600, seed=5)
    hyp = synth.make_hypotheses(gt, 64, seed=6, random_frac=0.1)
    torch.cuda.set_device(0)
    c0, c1 = hop_b200.Context(0), hop_b200.Context(1)
    out = []
    for c in (c1, c0, c1):   # device 1 first: nothing may depend on which device ran a kernel first
        scene, model = c.upload_cloud(s, sn, conf), c.upload_cloud(m, mn)
        got, it, cv = c.icp_refine(scene, model, hyp)
        sc = c.lcp_score(scene, model, got)
        top = c.select_topk(got, sc, 8)        # (the winners kernel needs its > 48 KB shared-memory opt-in on BOTH devices)
        out.append((got, it, cv, sc, top))
        scene.free(); model.free()
        assert torch.cuda.current_device() == 0
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and np.array_equal(o[1], out[0][1]) and np.array_equal(o[3], out[0][3])
        assert np.array_equal(o[4]["id"], out[0][4]["id"])
    c0.close(); c1.close()


def _rank(rank, world, id_path, H, K, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
    import time
    import hop_b200
    from hop_b200 import capi, distributed as D
    if rank == 0:
        uid = capi.comm_unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_path + ".tmp", id_path)
    else:
        for _ in range(600):
            if os.path.exists(id_path):
                break
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    ctx = hop_b200.Context(rank)
    ctx.comm_init(uid, rank, world)
    m, mn = synth.make_model("ellipse", 3000, seed=1)
    s, sn, conf, gt = synth.make_scene("ellipse", 600, seed=5)
    hyp = synth.make_hypotheses(gt, H, seed=6, random_frac=0.1)
    scene, model = ctx.upload_cloud(s, sn, conf), ctx.upload_cloud(m, mn)
    b, e = D.shard_range(H, rank, world)
    got, sc, it, cv, win = ctx.refine_score_select(scene, model, hyp[b:e], K=K)
    win["id"][win["id"] >= 0] += b
    allrec = ctx.gather_winners(win)                       # host-buffer entry point
    merged = D.merge_winners(allrec, K)
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    if rank == 0:                                          # the single-GPU answer on the whole batch
        g, s1, _, _, w1 = ctx.refine_score_select(scene, model, hyp, K=K)
        np.save(os.path.join(out_dir, "single.npy"), w1)
    ctx.close()


@needs2
def test_two_ranks_gather_winners_through_the_c_abi(tmp_path):
    import torch.multiprocessing as mp
    H, K, world = 301, 8, 2
    mp.spawn(_rank, args=(world, str(tmp_path / "nccl_id"), H, K, str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npy")
    for r in range(world):
        mrg = np.load(tmp_path / f"merged_{r}.npy")
        assert np.array_equal(mrg["id"], single["id"]) and np.array_equal(mrg["score"], single["score"])
        assert np.array_equal(mrg["pose"], single["pose"])


def test_single_rank_gather_is_a_copy(ctx):
    from hop_b200 import capi
    rec = np.zeros(4, capi.POSE_REC_DTYPE)
    rec["id"] = [3, 1, 2, -1]; rec["score"] = [5, 4, 3, -np.inf]
    out = ctx.gather_winners(rec)
    assert np.array_equal(out, rec)
