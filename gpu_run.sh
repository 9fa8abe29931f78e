export B2D_EXPERIMENTAL=1
B2D_FIR_AUTO=fft timeout 45 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_fft_auto.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fft_auto.log; tail -4 gpurun_out/pytest_fft_auto.log
timeout 25 python bench.py --fir-impl fft --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_fir_fft.json 2> gpurun_out/b_fir_fft.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/b_fir_fft.json").read().strip().splitlines()[-1])
    print("fft", round(d["value"]), round(d["ms_per_step"],4), d["roofline"]["kernel_ms"], "e2e", round(d["e2e"]["value"]))
except Exception as e: print("ERR", e, open("gpurun_out/b_fir_fft.err").read()[-600:])
PY
