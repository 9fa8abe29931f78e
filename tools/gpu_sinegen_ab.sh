#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_combsub_sinegen.py -q -k "sinegen or source_module" > gpurun_out/pytest_sinegen_r7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sinegen_r7.log; tail -4 gpurun_out/pytest_sinegen_r7.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/parity_report.json"))
for k,v in d.items():
    if k.startswith("sinegen_noise"): print(k, {a:(float('%.3g'%b) if isinstance(b,float) else b) for a,b in v.items()})
PY
for impl in v2 v2r7 v2p; do for wl in sinegen srcmod; do
  timeout 100 python bench.py --workload $wl --sinegen-impl $impl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_${wl}_$impl.json 2> gpurun_out/b_${wl}_$impl.err
  python -c "
import json; d=json.loads(open('gpurun_out/b_${wl}_$impl.json').read().strip().splitlines()[-1]); print('$wl $impl', round(d['ms_per_step'],4), round(d['value']), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()}, round(d['roofline']['frac'],3))"
done; done
