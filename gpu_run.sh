timeout 300 python -m pytest tests/test_gpu_combsub_sinegen.py -q -m gpu -k "sinegen or source_module" 2>&1 | tail -4
for impl in v2p v2p8; do
  for wl in sinegen srcmod; do
    timeout 200 python bench.py --workload $wl --sinegen-impl $impl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_${wl}_${impl}.json 2>/dev/null
  done
done
python - <<'PY'
import json
for impl in ("v2p","v2p8"):
    for wl in ("sinegen","srcmod"):
        try:
            d=json.loads(open("gpurun_out/b_%s_%s.json"%(wl,impl)).read().strip().splitlines()[-1])
            print(wl, impl, round(d["value"]), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), d["roofline"]["kernel_ms"], round(d["e2e"]["value"]))
        except Exception as e: print(wl, impl, "ERR", e)
PY
