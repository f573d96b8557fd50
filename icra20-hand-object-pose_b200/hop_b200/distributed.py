"""Multi-GPU plumbing of the path: hypotheses / frames are independent units, so ranks work on disjoint
shards with no data-path collective, and ONE all-gather of fixed-size winner records (hop_pose_rec, 80 bytes, mirrors
class PoseHypo) ends a step.  Every rank then merges the same gathered list, so the result is replicated.

torch.distributed is used for the collective only (NCCL on GPUs -- the send buffer is written in place by
hop_select_topk_dev --, gloo in the CPU tests).
"""
import numpy as np

from .capi import POSE_REC_DTYPE


def shard_range(n, rank, world):
    """[begin, end) of the contiguous slice of n units owned by `rank`: the first n % world ranks take one extra unit."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def frames_for_rank(n_frames, rank, world):
    """frame f belongs to rank f % world (C4: 128 frames over 8 GPUs = 16 frames per rank)."""
    return list(range(rank, n_frames, world))


def empty_records(k):
    rec = np.zeros(k, POSE_REC_DTYPE)
    rec["id"] = -1
    rec["score"] = -np.inf
    rec["pose"][:, [0, 5, 10, 15]] = 1.0
    return rec


def local_winners(poses_colmajor, scores, k, id_offset=0, frame=0):
    """Host restatement of hop_select_topk's contract (score desc, ties -> lower id; unused slots id = -1, score = -inf)."""
    scores = np.asarray(scores, np.float32)
    s = np.where(np.isnan(scores), -np.inf, scores)
    order = np.lexsort((np.arange(len(s)), -s))[:k]
    rec = empty_records(k)
    rec["pose"][: len(order)] = np.asarray(poses_colmajor, np.float32).reshape(-1, 16)[order]
    rec["score"][: len(order)] = scores[order]
    rec["id"][: len(order)] = order + id_offset
    rec["frame"] = frame
    return rec


def gather_winners(local, group=None):
    """The path's one collective: all-gather of every rank's K records.  `local`: numpy record array (CPU / gloo) or a
    uint8 CUDA tensor of K * 80 bytes (the buffer hop_select_topk_dev wrote).  Returns the world x K records (numpy)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if isinstance(local, np.ndarray):
        send = torch.from_numpy(np.ascontiguousarray(local).view(np.uint8).copy())
    else:
        send = local
    if world == 1:
        out = send
    else:
        out = torch.empty(world * send.numel(), dtype=torch.uint8, device=send.device)
        dist.all_gather_into_tensor(out, send, group=group)
    return np.frombuffer(out.cpu().numpy().tobytes(), dtype=POSE_REC_DTYPE).copy()


def merge_winners(records, k):
    """Identical on every rank: the best k of the gathered records by (score desc, id asc), empty slots dropped."""
    rec = records[records["id"] >= 0]
    order = np.lexsort((rec["id"], -rec["score"]))[:k]
    return rec[order]
