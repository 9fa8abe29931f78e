"""Host logic of the multi-GPU path, on CPU with the gloo backend and world_size 2 (and 3 for
ragged shards): contiguous batch split, global utterance offsets, gather order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddsp_svc_b200 import sharding


def test_shard_bounds_cover_the_batch():
    for n in (1, 2, 7, 32, 256):
        for w in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = sharding.shard_sizes(n, w)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, chunks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = 64
        s, e = sharding.shard_bounds(n_global, world, rank)
        # each global utterance u is a waveform filled with u + t/1000: stands in for the synthesis,
        # which depends only on the GLOBAL utterance index (utterance_offset = s)
        rows = torch.arange(s, e, dtype=torch.float32)[:, None] + torch.arange(T, dtype=torch.float32)[None, :] / 1000
        out = sharding.gather_waveform(rows, n_global, dst=0, chunks=chunks)
        if rank == 0:
            want = torch.arange(n_global, dtype=torch.float32)[:, None] + torch.arange(T, dtype=torch.float32)[None, :] / 1000
            ret.put(bool(torch.equal(out, want)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global,chunks", [(2, 8, 1), (2, 7, 1), (3, 8, 2), (2, 8, 4)])
def test_gather_waveform_gloo(world, n_global, chunks):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, chunks, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True


def _worker_pipeline(rank, world, port, n_local, chunks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = 32
        calls = []

        def synth(lo, hi):
            calls.append((lo, hi))
            u = torch.arange(rank * n_local + lo, rank * n_local + hi, dtype=torch.float32)
            return u[:, None] * 10 + torch.arange(T, dtype=torch.float32)[None, :] / 100

        out = sharding.synthesize_and_gather(synth, n_local, n_local * world, T, "cpu", dst=0, chunks=chunks)
        ok = calls == sharding._chunk_bounds(n_local, chunks)
        if rank == 0:
            want = torch.arange(n_local * world, dtype=torch.float32)[:, None] * 10 + torch.arange(T, dtype=torch.float32)[None, :] / 100
            ret.put(bool(ok and torch.equal(out, want)))
        else:
            assert out is None and ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_local,chunks", [(2, 8, 4), (2, 5, 2), (3, 4, 1)])
def test_synthesize_and_gather_pipeline_gloo(world, n_local, chunks):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pipeline, args=(r, world, port, n_local, chunks, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True


def test_host_pipeline_chunk_bounds():
    """even and tapered utterance chunk schedules of the host-buffer pipeline cover the batch exactly once"""
    from ddsp_svc_b200.pipeline import chunk_bounds
    assert chunk_bounds(32, 4) == [(0, 8), (8, 16), (16, 24), (24, 32)]
    assert chunk_bounds(5, 4) == [(0, 2), (2, 3), (3, 4), (4, 5)]
    assert chunk_bounds(2, 8) == [(0, 1), (1, 2)]
    assert chunk_bounds(32, (4, 9, 13, 6)) == [(0, 4), (4, 13), (13, 26), (26, 32)]
    for batch in (1, 2, 3, 7, 32, 33, 64):
        for sched in ((4, 9, 13, 6), (1, 1), (4, 8, 12, 6, 2), (0, 1, 0)):
            b = chunk_bounds(batch, sched)
            assert b[0][0] == 0 and b[-1][1] == batch
            assert all(x[1] == y[0] for x, y in zip(b, b[1:])) and all(hi > lo for lo, hi in b)
    sizes = [hi - lo for lo, hi in chunk_bounds(64, (4, 9, 13, 6))]
    assert sizes == [8, 18, 26, 12]
    import pytest
    with pytest.raises(ValueError):
        chunk_bounds(4, (0, 0))
