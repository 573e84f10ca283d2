"""Multi-GPU plumbing: the path shards by independent unit (30 s window, 10 s diarization window, speech segment),
so ranks never exchange data on the hot path.  The one real exchange is the all-gather of per-rank speaker embeddings ahead of
global clustering.  On the GPU box that exchange runs INSIDE the library (csrc/dist.cu: wdr_dist_init / wdr_allgather_embeddings,
NCCL over NVLink on device buffers); `connect()` only carries the 128-byte NCCL id from rank 0 to the others over the launcher's
torch.distributed group, which is what a Rust host would do over its own channel.  `allgather_embeddings` below is the same
exchange over torch.distributed for the CPU-only (gloo) test of the host logic."""
import numpy as np


def connect(capi, device):
    """One wdr_dist per rank from an initialised torch.distributed group (any backend): rank 0 draws the NCCL id through the C ABI,
    the group broadcasts the bytes, every rank calls wdr_dist_init."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    box = [capi.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return capi.Dist(box[0], world, rank, device)


def shard_range(n_units, rank, world):
    """Contiguous block of unit ids for `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_embeddings(emb_local, device=None):
    """emb_local [N_r, D] (numpy or torch) on every rank -> [sum N_r, D] numpy in rank order (segments stay in time order
    when shards are contiguous).  Pads to the largest N_r; counts travel in a first tiny all-gather."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(emb_local, np.float32)
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    e = torch.as_tensor(np.asarray(emb_local, np.float32), device=device)
    n_local = torch.tensor([e.shape[0], e.shape[1] if e.ndim == 2 else 0], device=device, dtype=torch.int64)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    n_max = max(int(c[0]) for c in counts)
    D = max(int(c[1]) for c in counts)
    pad = torch.zeros(n_max, D, device=device, dtype=torch.float32)
    if e.numel():
        pad[: e.shape[0]] = e
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return np.concatenate([o[: int(c[0])].cpu().numpy() for o, c in zip(out, counts)], 0)
