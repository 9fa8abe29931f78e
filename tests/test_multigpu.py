"""Two-GPU test of the peer-mapped gather (needs >= 2 CUDA devices; skipped otherwise): every
rank's Sins forward writes `signal` straight into rank 0's symmetric-memory buffer and the result
must equal the single-GPU synthesis of the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from ddsp_svc_b200 import FixedControls, Sins, sharding, synthetic as syn
        SR, P, nF, H, Bl = 44100, 512, 20, 32, 3
        sm = syn.sins_split_map(H, 256, 256)
        f0 = syn.make_f0(world * Bl, nF)
        dense = syn.make_ctrl(world * Bl, nF, sm)[0]
        lo, hi = sharding.shard_bounds(world * Bl, world, rank)
        fixed = FixedControls(syn.split_views(dense[lo:hi].to(dev), sm), None)
        model = Sins(SR, P, H, 256, 256, unit2ctrl=fixed).to(dev)
        peer = sharding.PeerGather(Bl, nF * P, dev, dst=0)
        torch.manual_seed(11)            # same host seed on every rank -> same Philox key; rows differ by utterance index
        with torch.no_grad():
            sig, _, _ = model(None, f0[lo:hi].to(dev), None, utterance_offset=lo, signal_out=peer.my_rows)
            out = peer.finish()
            torch.cuda.synchronize()
            if rank == 0:
                fixed.ctrls = syn.split_views(dense.to(dev), sm)
                torch.manual_seed(11)
                ref, _, _ = model(None, f0.to(dev), None, utterance_offset=0)
                ret.put(float((out - ref).abs().max().item()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_gather_two_gpus():
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get() == 0.0
