"""Multi-GPU plumbing for the hot path: one process per GPU, independent units (polynomials of a
batch, whole commits) sharded across ranks, no data-path collective.  The only exchange is the
all-gather of per-shard commitments / evaluations into every rank's host Fiat-Shamir transcript
(tens of KiB, latency bound) -- NCCL on GPUs, gloo in the CPU tests."""
import hashlib

import numpy as np


def shard_range(n_units, rank, world):
    """Contiguous block partition of n_units independent units; returns (start, stop)."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_commitments(local, n_total, rank, world, device=None):
    """local: uint64 array [n_local, words] of this rank's commitments (units shard_range(rank)).
    Returns the [n_total, words] array in unit order on every rank."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.uint64)
    words = local.shape[1] if local.ndim == 2 else 0
    if world == 1:
        return local.copy()
    counts = [shard_range(n_total, r, world) for r in range(world)]
    max_n = max(b - a for a, b in counts)
    pad = np.zeros((max_n, words), dtype=np.uint64)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad.view(np.int64))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = []
    for r, (a, b) in enumerate(counts):
        parts.append(outs[r].cpu().numpy().view(np.uint64)[: b - a])
    return np.concatenate(parts, axis=0)


def transcript_digest(commitments):
    """SHA-256 over the gathered commitments in unit order (what buckler binds, prover.go:155-157)."""
    return hashlib.sha256(np.ascontiguousarray(commitments, dtype="<u8").tobytes()).hexdigest()
