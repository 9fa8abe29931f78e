#!/usr/bin/env bash
# A/B of the bank kernel's double-angle anchor chains (compile-time variant in libb200ddsp_da.so)
set -u
mkdir -p gpurun_out
D=$PWD/ddsp_svc_b200
B2D_LIB_PATH=$D/libb200ddsp_da.so timeout 300 python -m pytest tests/test_gpu_sins.py -q > gpurun_out/pytest_bank_da.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bank_da.log; tail -3 gpurun_out/pytest_bank_da.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/parity_report.json"))
for k,v in d.items():
    if k.startswith("stages/") or k in ("sins_truth","full_size"): print(k, {a:(float('%.3g'%b) if isinstance(b,float) else b) for a,b in v.items() if 'bank' in a or 'rms' in a or 'truth' in a})
PY
for lib in libb200ddsp.so libb200ddsp_da.so; do
  B2D_LIB_PATH=$D/$lib timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-others --breakdown > gpurun_out/b_ab_$lib.json 2> gpurun_out/b_ab_$lib.err
  echo "$lib $(tail -1 gpurun_out/b_ab_$lib.err)"; python -c "
import json; d=json.loads(open('gpurun_out/b_ab_$lib.json').read().strip().splitlines()[-1]); print('   step', round(d['ms_per_step'],4))"
done
