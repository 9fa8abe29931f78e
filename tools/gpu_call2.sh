#!/usr/bin/env bash
# round 2, call 2: GPU tests with today's default dispatch, overlap-mode sweep x register-cap variants of the IR kernel,
# ncu of the bank / IR kernels.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log; tail -3 gpurun_out/pytest_gpu_r2a.log
q() { name=$1; lib=$2; shift 2; B2D_LIB_PATH=$lib timeout 120 python bench.py --quick --steps 20 --warmup 3 "$@" > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err; echo "$name $(cat gpurun_out/q_$name.json | tail -1)"; }
D=$PWD/ddsp_svc_b200
for m in 0 1 2 3 4 8 -2 -4; do q ov${m}_def $D/libb200ddsp.so --overlap $m; done
for m in 1 2 4 -2 -4; do q ov${m}_r64 $D/libb200ddsp_r64.so --overlap $m; done
for m in 1 2 4; do q ov${m}_r96 $D/libb200ddsp_r96.so --overlap $m; done
q ov4_def_packed $D/libb200ddsp.so --overlap 4 --fft-arith packed
# ncu: bank + both IR builds under the default dispatch (in-order mode so each kernel is alone on the GPU)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'sins_bank_kernel|ir_build_tc_kernel' -s 6 -c 3 -f -o gpurun_out/prof_r2_bank_ir \
    python bench.py --steps 2 --warmup 3 --quick --overlap 0 > /dev/null 2>&1
for k in sins_bank_kernel ir_build_tc_kernel; do python tools/ncu_summary.py gpurun_out/prof_r2_bank_ir.ncu-rep $k > gpurun_out/ncu_r2_$k.txt 2>&1; done
# launch list of one default step (shares)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --quick --overlap 0 > /dev/null 2>&1
