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


def block_range(n_units, rank, world):
    """Equal contiguous blocks of ceil(n / world) units (the last ranks may own fewer or none): the split of
    rs_jindo_shard_columns.  Laid end to end the blocks are the whole array, so ONE in-place all-gather of
    equal-sized slices exchanges them."""
    per = -(-n_units // world)
    return min(rank * per, n_units), min((rank + 1) * per, n_units)


class NcclComm:
    """An NCCL communicator owned by the library (rs_nccl_comm_create): rank 0 makes the unique id, the other ranks get
    it over the host's own control plane -- here a torch.distributed object broadcast, in a Go host any RPC."""

    def __init__(self, rank, world):
        import ctypes
        import torch.distributed as dist
        from .capi import check, lib, c_vp
        self.rank, self.world = rank, world
        ident = [None]
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            check(lib().rs_nccl_unique_id(ctypes.cast(buf, c_vp)))
            ident[0] = buf.raw
        dist.broadcast_object_list(ident, src=0)
        self.h = c_vp()
        check(lib().rs_nccl_comm_create(ctypes.c_char_p(ident[0]), rank, world, ctypes.byref(self.h)))

    def all_gather_u64(self, local):
        """local: uint64 array of n words -> [world, n] on every rank (rs_allgather_u64)"""
        import ctypes
        from .capi import check, lib
        local = np.ascontiguousarray(local, dtype=np.uint64).reshape(-1)
        out = np.zeros((self.world, local.size), dtype=np.uint64)
        check(lib().rs_allgather_u64(self.h, self.rank, self.world, local.ctypes.data_as(ctypes.c_void_p), local.size,
                                     out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def close(self):
        from .capi import lib
        if self.h:
            lib().rs_nccl_comm_destroy(self.h)
            self.h = None


def sharded_commit(prover, batch, poly, length, tape, rank, world, opening=None, comm=None, gather_incommit=True):
    """One Commit split over `world` GPUs by columns of the committed matrix (SURVEY 8e iii), the exchange inside the
    library: each rank encodes and inner-commits its block of columns (rs_jindo_commit_columns), multiplies ck.Out by its
    own InCommit polynomials, the out_msis partial sums are all-reduced over NCCL and the cut-off pipeline runs replicated
    (rs_jindo_outer_commit_sharded) -- nothing but a few KiB is on Commit's critical path.  The InCommit slices that
    Evaluate / the proof need are all-gathered in place afterwards (rs_jindo_allgather_incommit, asynchronous).
    Returns (commitment, opening); Encode / MLWE stay sharded by column."""
    import ctypes
    from .capi import check, lib
    if world == 1:
        return prover.Commit(batch, poly, length, tape)
    c0, c1 = ctypes.c_int(), ctypes.c_int()
    check(lib().rs_jindo_shard_columns(prover.h, rank, world, ctypes.byref(c0), ctypes.byref(c1)))
    if c1.value > c0.value:
        opening = prover.CommitColumns(batch, poly, length, tape, c0.value, c1.value, opening)
    elif opening is None:
        opening = prover.NewOpening()
    com = np.zeros((prover.p["out_msis"], prover.n_qout, prover.d), dtype=np.uint64)
    check(lib().rs_jindo_outer_commit_sharded(prover.h, opening.h, rank, world, comm.h, com.ctypes.data_as(ctypes.c_void_p)))
    if gather_incommit:
        check(lib().rs_jindo_allgather_incommit(prover.h, opening.h, rank, world, comm.h))
    return com, opening


def transcript_digest(commitments):
    """SHA-256 over the gathered commitments in unit order (what buckler binds, prover.go:155-157)."""
    return hashlib.sha256(np.ascontiguousarray(commitments, dtype="<u8").tobytes()).hexdigest()
