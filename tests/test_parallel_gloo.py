"""CPU test of the N>1 path: world_size-2 gloo processes shard independent units, all-gather their
(synthetic) commitments and must end with the same transcript on every rank, identical to the
single-process order."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _unit_commitment(u, words):
    rng = np.random.default_rng(1000 + u)
    return rng.integers(0, 1 << 62, size=words, dtype=np.uint64)


def _worker(rank, world, n_total, words, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from conftest import load_product
    par = load_product().parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = par.shard_range(n_total, rank, world)
    local = np.stack([_unit_commitment(u, words) for u in range(a, b)]) if b > a else np.zeros((0, words), dtype=np.uint64)
    full = par.all_gather_commitments(local, n_total, rank, world)
    q.put((rank, (a, b), par.transcript_digest(full)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 2, 1])
def test_sharded_commitments_allgather(n_total, product):
    world, words = 2, 6 * 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_total) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, n_total, words, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = product.parallel.transcript_digest(np.stack([_unit_commitment(u, words) for u in range(n_total)]))
    ranges = sorted(r[1] for r in res)
    assert ranges[0][0] == 0 and ranges[-1][1] == n_total and ranges[0][1] == ranges[1][0]
    assert all(r[2] == expect for r in res)


def test_shard_range_partition(product):
    par = product.parallel
    for n in (0, 1, 5, 8, 513):
        for w in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _slice_worker(rank, world, n_cols, words, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_product
    par = load_product().parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_cols * words, dtype=torch.int64) * 7 + 3
    buf = torch.zeros_like(full)
    a, b = par.shard_range(n_cols, rank, world)
    buf[a * words:b * words] = full[a * words:b * words]          # this rank's columns only
    par.gather_column_slices(buf, n_cols, words, world)
    q.put((rank, bool(torch.equal(buf, full))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [129, 3, 1])
def test_column_slices_exchange(n_cols, product):
    """The InCommit exchange of a column-sharded Commit (parallel.gather_column_slices), world 2 on gloo."""
    world, words = 2, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + n_cols) % 2000
    procs = [ctx.Process(target=_slice_worker, args=(r, world, n_cols, words, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_block_range_is_the_padded_allgather_layout(product):
    """rs_jindo_shard_columns' split: equal blocks of ceil(n / world); rank r's block starts at r * per, so equal-sized
    slots laid end to end reproduce the array (what makes the exchange ONE in-place all-gather)."""
    par = product.parallel
    for n in (129, 5, 1, 9, 64):
        for w in (1, 2, 3, 4, 8, 16):
            per = -(-n // w)
            spans = [par.block_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and max(b for _, b in spans) == n
            assert all(a == min(r * per, n) for r, (a, _) in enumerate(spans))
            assert all(b - a == per for (a, b) in spans if b < n)             # every block before the end is full
            assert w * per >= n and w * per - n < w                           # slack the Opening's buffer must provide


def _block_worker(rank, world, n_cols, words, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_product
    par = load_product().parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = -(-n_cols // world)
    full = torch.arange(n_cols * words, dtype=torch.int64) * 5 + 1
    buf = torch.zeros(world * per * words, dtype=torch.int64)              # the Opening's InCommit buffer with its slack
    a, b = par.block_range(n_cols, rank, world)
    buf[a * words:b * words] = full[a * words:b * words]
    # the library's exchange: all-gather of slot `rank` (in place on NCCL; a list of views here, gloo has no in-place form)
    slots = [buf[r * per * words:(r + 1) * per * words] for r in range(world)]
    dist.all_gather(slots, slots[rank].clone())
    q.put((rank, bool(torch.equal(buf[:n_cols * words], full))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [129, 2, 5])
def test_padded_block_allgather(n_cols, product):
    world, words = 2, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + n_cols) % 2000
    procs = [ctx.Process(target=_block_worker, args=(r, world, n_cols, words, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
