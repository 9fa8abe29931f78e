#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_mel.py tests/test_gpu_combsubfast.py -q > gpurun_out/pytest_f2f4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_f2f4.log; tail -25 gpurun_out/pytest_f2f4.log
for wl in mel maskmul; do timeout 120 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_$wl.json 2> gpurun_out/b_$wl.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_$wl.json").read().strip().splitlines()[-1]); print("$wl", round(d["value"]), round(d["ms_per_step"],4), d["roofline"]["kernel_ms"], round(d["roofline"]["frac"],3), round(d["e2e"]["ms_per_step"],3))
except Exception as e: print("$wl ERR", e, open("gpurun_out/b_$wl.err").read()[-800:])
PY
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_mel python bench.py --workload mel --steps 3 --warmup 3 --quick > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2_mel.ncu-rep mel_kernel > gpurun_out/ncu_r2_mel.txt 2>&1; head -24 gpurun_out/ncu_r2_mel.txt
