#!/usr/bin/env bash
# Round-2 evidence run on one B200: smoke, full GPU suite, the default bench line (incl. other_workloads, eager-GPU and CPU
# baselines), the reference arm, the ncu launch list of one step and full captures of the Sins kernels.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -1 gpurun_out/smoke_r2.log
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_r2_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2_final.log; tail -3 gpurun_out/pytest_gpu_r2_final.log
cp gpurun_out/parity_report.json gpurun_out/parity_report_r2_final.json 2>/dev/null
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; tail -c 300 gpurun_out/bench_r2_default.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; tail -c 400 gpurun_out/bench_r2_reference.json; echo
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --quick --overlap 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'ltv_fir_fft_kernel|sins_bank_kernel|ir_build_tc_kernel' -s 8 -c 4 -f -o gpurun_out/prof_r2_sins_final \
    python bench.py --steps 2 --warmup 3 --quick --overlap 0 > /dev/null 2>&1
for k in ltv_fir_fft_kernel sins_bank_kernel ir_build_tc_kernel; do python tools/ncu_summary.py gpurun_out/prof_r2_sins_final.ncu-rep $k > gpurun_out/ncu_r2_final_$k.txt 2>&1; done
head -8 gpurun_out/ncu_r2_final_ltv_fir_fft_kernel.txt
