"""The N > 1 host logic on CPU: two gloo ranks shard a hypothesis batch, exchange their winner records with the path's one
all-gather and merge them -- every rank must end with the single-process answer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, K, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
    import torch.distributed as dist
    from hop_b200 import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                      # the same batch on every rank; each scores only its shard
    poses = rng.normal(size=(H, 16)).astype(np.float32)
    scores = rng.integers(0, 30, H).astype(np.float32)    # many ties: the id tie-break must survive the exchange
    b, e = D.shard_range(H, rank, world)
    local = D.local_winners(poses[b:e], scores[b:e], K, id_offset=b, frame=rank)
    merged = D.merge_winners(D.gather_winners(local), K)
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    dist.destroy_process_group()


@pytest.mark.parametrize("H,K", [(1001, 16), (7, 16)])
def test_two_ranks_agree_with_one(tmp_path, H, K):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
    from hop_b200 import distributed as D
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, H, K, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    poses = rng.normal(size=(H, 16)).astype(np.float32)
    scores = rng.integers(0, 30, H).astype(np.float32)
    single = D.merge_winners(D.local_winners(poses, scores, K), K)
    for r in range(world):
        m = np.load(os.path.join(str(tmp_path), f"merged_{r}.npy"))
        assert np.array_equal(m["id"], single["id"]) and np.array_equal(m["score"], single["score"])
        assert np.array_equal(m["pose"], single["pose"])
    assert len(single) == min(H, K)


def test_shards_partition_the_batch():
    sys.path.insert(0, os.path.join(ROOT, "icra20-hand-object-pose_b200"))
    from hop_b200 import distributed as D
    for n in (0, 1, 7, 1024, 4097):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    assert D.frames_for_rank(128, 3, 8) == list(range(3, 128, 8)) and len(D.frames_for_rank(128, 0, 8)) == 16
