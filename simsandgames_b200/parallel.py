"""Host-side plumbing for the slab-sharded path (one process per GPU, launched by torchrun).

torch.distributed is used only to hand the NCCL communicator id from rank 0 to the other ranks and to gather
results; the data path (halo rows) goes through the library's own NCCL communicator over NVLink."""
import numpy as np


def broadcast_unique_id(make_id=None):
    """Rank 0 creates the 128-byte id (default: the library's ncclGetUniqueId), every rank returns the same bytes."""
    import torch
    import torch.distributed as dist
    if make_id is None:
        from .binding import nccl_unique_id as make_id
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(make_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def slab_partition(ns, nranks):
    """[(j0, j1)] for every rank: contiguous, exhaustive, the library's rule (sorshard_slab)."""
    return [(ns * r // nranks, ns * (r + 1) // nranks) for r in range(nranks)]


def stitch_rows(parts):
    """Concatenate per-rank own-row blocks (in rank order) back into the global (ns, nf) array."""
    return np.concatenate(parts, axis=0)


def weighted_partition(row_weights, nranks, min_rows=10):
    """Row boundaries [r_0=0, ..., r_nranks=ns] that give every rank about the same total weight (a rank's work = its rows'
    weights: particles x cost per particle + cells x cost per cell), with at least min_rows rows per slab (halo + 1)."""
    w = np.asarray(row_weights, np.float64)
    ns = w.size
    assert ns >= nranks * min_rows
    c = np.concatenate([[0.0], np.cumsum(w)])
    starts = [0]
    for r in range(1, nranks):
        j = int(np.searchsorted(c, c[-1] * r / nranks))
        j = max(j, starts[-1] + min_rows)
        j = min(j, ns - (nranks - r) * min_rows)
        starts.append(j)
    starts.append(ns)
    return starts
