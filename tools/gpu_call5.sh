#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2b.log; tail -15 gpurun_out/pytest_gpu_r2b.log
cp gpurun_out/parity_report.json gpurun_out/parity_report_r2b.json 2>/dev/null
B2D_FFT_ARITH=packed timeout 400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r2b_packed.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2b_packed.log; tail -5 gpurun_out/pytest_gpu_r2b_packed.log
q() { name=$1; shift; timeout 120 python bench.py --quick --steps 20 --warmup 3 "$@" > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err; echo "$name $(tail -1 gpurun_out/q_$name.json)"; }
q cfg1_ov1 --workload sins_cfg1
q cfg1_ov0 --workload sins_cfg1 --overlap 0
q sins_ov1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 600 gpurun_out/bench_r2b.json
