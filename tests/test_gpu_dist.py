"""GPU tests of the path's one exchange behind the C ABI (csrc/dist.cu): wdr_dist_init / wdr_allgather_embeddings[_dev] and
the speaker policy on the gathered table.  One rank needs no NCCL (the gather is a staged copy); the 2-rank test spawns one process per
GPU and runs only where two devices are visible (`gpurun --gpus 2`); the host logic of the N > 1 path is covered on CPU by
tests/test_dist_gloo.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _table(n, d=32, seed=5):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((4, d))
    return (cent[rng.integers(0, 4, n)] + 0.15 * rng.standard_normal((n, d))).astype(np.float32)


def test_single_rank_gather_normalises_and_policy_equals_leader_scan(wdr):
    from oracle import cluster as K
    E = _table(57)
    E[11] = 0.0  # a zero row stays zero
    d = wdr.Dist(None, 1, 0, 0)
    out, counts = d.allgather_embeddings(E, 100, normalize=True)
    assert out.shape == E.shape and list(counts) == [57]
    nrm = np.linalg.norm(E, axis=1, keepdims=True)
    ref = np.divide(E, nrm, out=np.zeros_like(E), where=nrm > 0)
    assert np.abs(out - ref).max() < 1e-6 and not out[11].any()
    raw, _ = d.allgather_embeddings(E, 57, normalize=False)
    assert np.array_equal(raw, E)
    with pytest.raises(wdr.WdrError):
        d.allgather_embeddings(E, 10)  # output too small
    d.close()
    # the crate's policy over the table == the ordered leader scan on its cosine matrix (oracle), also under a speaker cap
    for cap in (wdr.SIZE_MAX, 2):
        m = wdr.EmbeddingManager(cap)
        lab = m.assign_batch(E, 0.5)
        m.close()
        assert lab.tolist() == K.leader_labels(K.cosine_matrix(E), 0.5, cap if cap != wdr.SIZE_MAX else 10**9).tolist()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    import wdr_b200 as w
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # only carries the NCCL id: the exchange is the library's
    wd = w.dist.connect(w, rank)
    E = _table(41)
    lo, hi = w.dist.shard_range(len(E), rank, world)
    if rank == 1:
        hi = lo  # a rank with no embeddable segment
    host, counts = wd.allgather_embeddings(E[lo:hi], 64)
    loc = torch.from_numpy(np.ascontiguousarray(E[lo:hi])).cuda()
    outd = torch.zeros(64, E.shape[1], device="cuda")
    n, counts_d = wd.allgather_embeddings_dev(loc.data_ptr() if hi > lo else 0, hi - lo, E.shape[1], outd.data_ptr(), 64, normalize=True)
    q.put((rank, lo, hi, host.tolist(), counts.tolist(), int(n), outd[:n].cpu().numpy().tolist(), counts_d.tolist()))
    wd.close()
    dist.destroy_process_group()


def test_two_ranks_nccl_allgather(wdr):
    import torch
    if torch.cuda.device_count() < 2 or not wdr.dist_available():
        pytest.skip("needs two GPUs and NCCL (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = sorted(q.get(timeout=240) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:  # a rank that died leaves its peer inside a collective: never leave it on the GPU
            if p.is_alive():
                p.kill()
    E = _table(41)
    lo0, hi0 = res[0][1], res[0][2]
    want = E[lo0:hi0]  # rank 1 contributed nothing
    for r in res:
        assert r[4] == [hi0 - lo0, 0] and r[7] == [hi0 - lo0, 0] and r[5] == hi0 - lo0
        assert np.array_equal(np.asarray(r[3], np.float32), want)
        nrm = want / np.linalg.norm(want, axis=1, keepdims=True)
        assert np.abs(np.asarray(r[6], np.float32) - nrm).max() < 1e-6
