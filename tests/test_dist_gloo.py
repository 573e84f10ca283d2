"""World-size-2 gloo test of the N>1 path's host logic: contiguous shard ranges and the embedding all-gather that
precedes global clustering.  Runs on CPU; the GPU box uses the same code over NCCL."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    import wdr_b200 as w
    from oracle import cluster as K
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(77)  # same on every rank: the global embedding table
    cent = rng.standard_normal((3, 16))
    E = (cent[rng.integers(0, 3, n_total)] + 0.2 * rng.standard_normal((n_total, 16))).astype(np.float32)
    lo, hi = w.dist.shard_range(n_total, rank, world)
    gathered = w.dist.allgather_embeddings(E[lo:hi])
    labels = K.leader_labels(K.cosine_matrix(gathered), 0.5, 10**9)
    q.put((rank, lo, hi, bool(np.array_equal(gathered, E)), labels.tolist()))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    import wdr_b200 as w
    for n in (0, 1, 7, 120, 121):
        for world in (1, 2, 3, 8):
            r = [w.dist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_allgather_embeddings_world2_gloo():
    from oracle import cluster as K
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_total = 23  # ragged: 12 + 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 12), (12, 23)]
    assert all(r[3] for r in res)          # every rank reconstructed the global table in order
    assert res[0][4] == res[1][4]          # hence identical global labels on every rank
    rng = np.random.default_rng(77)
    cent = rng.standard_normal((3, 16))
    E = (cent[rng.integers(0, 3, n_total)] + 0.2 * rng.standard_normal((n_total, 16))).astype(np.float32)
    assert res[0][4] == K.leader_labels(K.cosine_matrix(E), 0.5, 10**9).tolist()
