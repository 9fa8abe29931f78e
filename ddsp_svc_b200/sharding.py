"""Batch sharding of the synthesis path across the GPUs of one box.

Every utterance (batch row) is independent end to end (reference ddsp/vocoder.py:556-611 has no
cross-row term), so the path shards trivially: rank r synthesizes a contiguous slice of the
global batch with ``utterance_offset`` = first global row (the in-kernel Philox noise is keyed by
the GLOBAL utterance index, so results do not depend on the sharding), and the only collective is
the final gather of the waveform.  Host logic here is backend-agnostic (NCCL on GPUs, gloo in the
CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous split, remainder spread over the first ranks: -> (start, stop)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_items, world_size):
    return [shard_bounds(n_items, world_size, r)[1] - shard_bounds(n_items, world_size, r)[0]
            for r in range(world_size)]


def gather_waveform(local, n_global, dst=0, group=None, chunks=1):
    """Gather the per-rank waveforms [B_local, T] into [n_global, T] on rank ``dst`` (None
    elsewhere).  Shards may be ragged.  ``chunks`` > 1 splits the local rows into that many
    point-to-point messages so the transfer of finished rows can overlap the synthesis of the
    rest when the caller issues this on a side stream."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_global, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local.shape[0], sizes[rank]))
    T = local.shape[1]
    if len(set(sizes)) == 1 and chunks == 1:
        out = torch.empty(n_global, T, dtype=local.dtype, device=local.device) if rank == dst else None
        dist.gather(local.contiguous(), list(out.split(sizes[0])) if rank == dst else None, dst=dst, group=group)
        return out
    # ragged or chunked: batched point-to-point
    if rank == dst:
        out = torch.empty(n_global, T, dtype=local.dtype, device=local.device)
        ops = []
        for r in range(world):
            s, e = shard_bounds(n_global, world, r)
            if r == dst:
                out[s:e].copy_(local)
                continue
            for cs, ce in _chunk_bounds(e - s, chunks):
                ops.append(dist.P2POp(dist.irecv, out[s + cs:s + ce], r, group))
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        return out
    ops = [dist.P2POp(dist.isend, local[cs:ce].contiguous(), dst, group)
           for cs, ce in _chunk_bounds(local.shape[0], chunks)]
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()
    return None


def _chunk_bounds(n, chunks):
    chunks = max(1, min(chunks, n)) if n > 0 else 1
    return [shard_bounds(n, chunks, c) for c in range(chunks) if shard_bounds(n, chunks, c)[1] > shard_bounds(n, chunks, c)[0]]


def synthesize_and_gather(synth_chunk, n_local, n_global, n_samples, device, dst=0, group=None, chunks=4,
                          dtype=torch.float32):
    """Chunked synthesis with the gather of finished chunks overlapped on a side stream.

    ``synth_chunk(lo, hi)`` synthesizes local rows [lo, hi) and returns a [hi-lo, n_samples] tensor
    on the current stream.  While chunk c+1 is being synthesized, chunk c travels to rank ``dst``
    (point-to-point over NCCL / NVLink).  Requires equal shard sizes.  Returns the gathered
    [n_global, n_samples] tensor on ``dst`` (None elsewhere); the current stream waits for the
    transfers before returning.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if n_local * world != n_global:
        raise ValueError("synthesize_and_gather needs equal shards (%d x %d != %d)" % (n_local, world, n_global))
    cuda = torch.device(device).type == "cuda"
    out = torch.empty(n_global, n_samples, dtype=dtype, device=device) if rank == dst else None
    bounds = _chunk_bounds(n_local, chunks)
    main = torch.cuda.current_stream(device) if cuda else None
    comm = _comm_stream(device) if cuda else None
    keep, works = [], []
    for lo, hi in bounds:
        part = synth_chunk(lo, hi)
        if part.shape != (hi - lo, n_samples):
            raise ValueError("synth_chunk returned %s, expected %s" % (tuple(part.shape), (hi - lo, n_samples)))
        keep.append(part)
        if cuda:
            ev = torch.cuda.Event()
            ev.record(main)
            comm.wait_event(ev)
        ctx = torch.cuda.stream(comm) if cuda else _null_ctx()
        with ctx:
            if rank == dst:
                out[dst * n_local + lo:dst * n_local + hi].copy_(part, non_blocking=True)
                ops = [dist.P2POp(dist.irecv, out[r * n_local + lo:r * n_local + hi], r, group)
                       for r in range(world) if r != dst]
            else:
                ops = [dist.P2POp(dist.isend, part, dst, group)]
            if ops:
                works.extend(dist.batch_isend_irecv(ops))
    ctx = torch.cuda.stream(comm) if cuda else _null_ctx()
    with ctx:
        for w in works:
            w.wait()
    if cuda:
        main.wait_stream(comm)
    del keep
    return out


_comm_streams = {}


def _comm_stream(device):
    key = torch.device(device).index
    if key not in _comm_streams:
        _comm_streams[key] = torch.cuda.Stream(device=device)
    return _comm_streams[key]


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class PeerGather:
    """Gather without a gather: rank ``dst`` owns the [n_global, T] result in symmetric memory and every
    rank holds a tensor view of ITS rows of that buffer, mapped over NVLink (CUDA peer / fabric
    handles via torch symmetric memory).  Passing ``my_rows`` as the synthesizer's ``signal_out`` makes
    the last kernel of the path store the waveform directly into rank ``dst``'s HBM, so the transfer
    overlaps the whole kernel instead of following it.  ``finish()`` is a device-side barrier over the
    group (stream ordered): after it, ``result`` on ``dst`` holds every rank's rows.
    """

    def __init__(self, n_local, n_samples, device, dst=0, group=None, dtype=torch.float32):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.dst, self.rank = dst, rank
        # symmetric allocation: every rank allocates the same shape; only dst's copy is the destination
        self._local = symm_mem.empty(world * n_local, n_samples, dtype=dtype, device=device)
        self._hdl = symm_mem.rendezvous(self._local, self.group)
        self.my_rows = self._hdl.get_buffer(dst, (n_local, n_samples), dtype, rank * n_local * n_samples)
        self.result = self._local if rank == dst else None

    def finish(self):
        self._hdl.barrier()
        return self.result

    def push_rows(self, lo, hi, rows, stream):
        """Copy finished local rows [lo, hi) into rank dst's buffer with the copy engine (DMA over NVLink)
        on ``stream``; no SM takes part, so it overlaps the synthesis of the next chunk."""
        with torch.cuda.stream(stream):
            self.my_rows[lo:hi].copy_(rows, non_blocking=True)
            rows.record_stream(stream)


def synthesize_and_push(synth_chunk, peer, n_local, device, chunks=4, streams=1, direct=False):
    """Chunked synthesis with the transfer of finished chunks to rank dst overlapped with the synthesis of the rest.

    ``synth_chunk(lo, hi, signal_out=None)`` synthesizes local rows [lo, hi) on the current stream and returns the
    [hi-lo, n_samples] waveform (written into ``signal_out`` when given).

    * ``direct=False``: each finished chunk is DMA-copied (copy engine, no SM) into rank dst's peer-mapped buffer on a
      side stream while the next chunk computes;
    * ``direct=True``: the chunk's last kernel stores the waveform straight into rank dst's buffer (peer stores over
      NVLink); with ``streams`` > 1 the back-pressured stores of chunk c overlap the arithmetic of chunk c+1.
    * ``streams`` > 1: consecutive chunks run on alternating compute streams, so chunk c+1's first kernels fill the SMs
      that chunk c's last wave leaves idle (and waveforms leave the GPU spread over the step, not at its end).

    Returns the gathered tensor on dst after a device-side barrier (stream ordered on the current stream)."""
    main = torch.cuda.current_stream(device)
    comm = _comm_stream(device)
    comm.wait_stream(main)
    side = _compute_streams(device, streams) if streams > 1 else []
    for s in side:
        s.wait_stream(main)
    for c, (lo, hi) in enumerate(_chunk_bounds(n_local, chunks)):
        cs = side[c % len(side)] if side else main
        with torch.cuda.stream(cs):
            if direct:
                part = synth_chunk(lo, hi, signal_out=peer.my_rows[lo:hi])
            else:
                part = synth_chunk(lo, hi)
                ev = torch.cuda.Event()
                ev.record(cs)
                comm.wait_event(ev)
                peer.push_rows(lo, hi, part, comm)
            del part
    for s in side:
        main.wait_stream(s)
    main.wait_stream(comm)
    return peer.finish()


_compute = {}


def _compute_streams(device, n):
    key = torch.device(device).index
    have = _compute.setdefault(key, [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device=device))
    return have[:n]
