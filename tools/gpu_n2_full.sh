#!/usr/bin/env bash
# the driver's N = 2 invocation of both arms, full (non-quick) lines, + the multi-GPU gather tests
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_multigpu.py -q -x > gpurun_out/pytest_multigpu_2_final.log 2>&1; tail -2 gpurun_out/pytest_multigpu_2_final.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_reference.json 2> gpurun_out/bench_n2_reference.err; tail -c 300 gpurun_out/bench_n2_reference.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_default.json 2> gpurun_out/bench_n2_default.err; tail -c 700 gpurun_out/bench_n2_default.json; echo; tail -3 gpurun_out/bench_n2_default.err
