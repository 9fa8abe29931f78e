#!/usr/bin/env bash
# Multi-GPU call: gather correctness on N GPUs (every mode), then a sweep of the gather modes with the quick bench.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_call_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/multi_gpus_$N.txt
timeout 400 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/pytest_multigpu_$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multigpu_$N.log; tail -4 gpurun_out/pytest_multigpu_$N.log
port=29500
run() { name=$1; shift; port=$((port+1));
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --quick --steps 20 --warmup 3 "$@" > gpurun_out/m${N}_$name.json 2> gpurun_out/m${N}_$name.err
  echo "$name $(tail -1 gpurun_out/m${N}_$name.json)"; }
if [ "$N" = "8" ]; then
run nogather --no-gather
run peer --gather peer
run pchunks2x2 --gather peer-chunks --gather-chunks 2 --gather-streams 2
run pchunks4x2 --gather peer-chunks --gather-chunks 4 --gather-streams 2
run pcopy2x1 --gather peer-copy --gather-chunks 2 --gather-streams 1
run pcopy2x2 --gather peer-copy --gather-chunks 2 --gather-streams 2
run pcopy4x2 --gather peer-copy --gather-chunks 4 --gather-streams 2
run nccl1 --gather nccl
else
run nogather --no-gather
run peer --gather peer
run pchunks2x2 --gather peer-chunks --gather-chunks 2 --gather-streams 2
run pchunks4x2 --gather peer-chunks --gather-chunks 4 --gather-streams 2
run pchunks4x1 --gather peer-chunks --gather-chunks 4 --gather-streams 1
run pcopy2x1 --gather peer-copy --gather-chunks 2 --gather-streams 1
run pcopy4x1 --gather peer-copy --gather-chunks 4 --gather-streams 1
run pcopy4x2 --gather peer-copy --gather-chunks 4 --gather-streams 2
run pcopy8x2 --gather peer-copy --gather-chunks 8 --gather-streams 2
run nccl1 --gather nccl
run nccl4 --gather nccl --gather-chunks 4
fi
