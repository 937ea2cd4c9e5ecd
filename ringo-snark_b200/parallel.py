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


class _DevView:
    """Minimal __cuda_array_interface__ holder: lets torch view library-owned device memory without a copy."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def gather_column_slices(buf, n_cols, words_per_col, world, group=None):
    """In-place exchange of a column-partitioned buffer: rank r owns columns shard_range(n_cols, r, world) of
    `buf` (a flat torch tensor of n_cols*words_per_col words, on any device); afterwards every rank holds all of
    them.  One broadcast per owner: the slices are a few MB (SURVEY 8e iii: 2.6 MB of InCommit at 2^22)."""
    import torch.distributed as dist
    for r in range(world):
        a, b = shard_range(n_cols, r, world)
        if b > a:
            dist.broadcast(buf[a * words_per_col:b * words_per_col], src=r, group=group)
    return buf


def sharded_commit(prover, batch, poly, length, tape, rank, world, opening=None):
    """One Commit split over `world` GPUs by columns of the committed matrix (SURVEY 8e iii): each rank encodes
    and inner-commits its columns, the InCommit slices are exchanged over NCCL directly in the Opening's device
    buffer, and every rank finishes with the same outer commit.  Returns (commitment, opening); the Opening
    stays sharded by column (Partial is per column; the responses need a modular all-reduce)."""
    import torch
    p = prover.p
    n_cols = p["cols"] + 1
    if world == 1:
        return prover.Commit(batch, poly, length, tape)
    c0, c1 = shard_range(n_cols, rank, world)
    opening = prover.CommitColumns(batch, poly, length, tape, c0, c1, opening)
    words = p["in_msis"] * prover.n_qout * prover.d
    view = torch.as_tensor(_DevView(opening.incommit_device_ptr(), n_cols * words), device="cuda")
    torch.cuda.synchronize()
    gather_column_slices(view, n_cols, words, world)
    torch.cuda.synchronize()
    return prover.OuterCommit(opening), opening


def transcript_digest(commitments):
    """SHA-256 over the gathered commitments in unit order (what buckler binds, prover.go:155-157)."""
    return hashlib.sha256(np.ascontiguousarray(commitments, dtype="<u8").tobytes()).hexdigest()
