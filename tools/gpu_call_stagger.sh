#!/usr/bin/env bash
# N = 8: does staggering the FIR sub-batches (library overlap mode k >= 2: lane 0 K0 F0 K2 F2 | side lane I K1 F1 K3 F3)
# spread the peer stores into rank 0 enough to hide them?
set -u
N=8
mkdir -p gpurun_out
port=29600
run() { name=$1; shift; port=$((port+1));
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --quick --steps 20 --warmup 3 "$@" > gpurun_out/s${N}_$name.json 2> gpurun_out/s${N}_$name.err
  echo "$name $(tail -1 gpurun_out/s${N}_$name.json)"; }
run peer --gather peer
run peer_ov2 --gather peer --overlap 2
run peer_ov4 --gather peer --overlap 4
run peer_ovm4 --gather peer --overlap -4
