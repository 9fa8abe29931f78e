#!/usr/bin/env bash
# quick check after a kernel change: the GPU tests of the Sins path + device-timed Sins / cfg1 steps + per-kernel times
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ir_tc.py tests/test_gpu_sins.py tests/test_gpu_fir_fft.py tests/test_gpu_combsub_sinegen.py tests/test_gpu_acceptance.py -q -x -k "not sinegen" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log; tail -4 gpurun_out/pytest_quick.log
q() { name=$1; shift; timeout 120 python bench.py --quick --steps 20 --warmup 3 "$@" > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err; echo "$name $(tail -1 gpurun_out/q_$name.json)"; }
q sins
q sins_split --sins-impl split
q sins_spectrum --sins-impl spectrum
q cfg1 --workload sins_cfg1
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-others --breakdown > gpurun_out/b_quick.json 2> gpurun_out/b_quick.err; tail -2 gpurun_out/b_quick.err
