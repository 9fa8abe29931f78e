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
