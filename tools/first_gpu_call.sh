#!/usr/bin/env bash
# The measurements that were still open when the round-1 GPU budget ran out, in priority order (≈ 3 min of box time).
# Run on a B200 from the repo root:   gpurun --timeout 600 -- 'bash tools/first_gpu_call.sh'
# Everything lands in gpurun_out/; copy what is worth keeping into profiles/.
set -u
mkdir -p gpurun_out
export B2D_EXPERIMENTAL=1            # also run the variants that have not executed on hardware yet
timeout 120 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_experimental.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_experimental.log; tail -3 gpurun_out/pytest_gpu_experimental.log
unset B2D_EXPERIMENTAL
b() { name=$1; shift; timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${NO_OTHERS:+--no-others} "$@" > gpurun_out/b_$name.json 2> gpurun_out/b_$name.err; }
b sins_default
export NO_OTHERS=1
b sins_packed --fft-arith packed
b sins_direct --fir-impl cuda
b sins_2streams --e2e-streams 2
b sins_even4 --e2e-chunks 4
b combsub --workload combsub
b combsubfast --workload combsubfast
b superfast --workload superfast
b superfast_packed --workload superfast --fft-arith packed
b cfg1 --workload sins_cfg1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/b_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-32s %9.0f Msamples/s  %.4f ms  e2e %8.0f (%.3f ms)  %s" % (
            f[13:-5], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
            {k: round(v, 3) for k, v in d["roofline"]["kernel_ms"].items()}))
    except Exception as e:
        print(f, "ERR", e)
PY
# ncu of the new dominant kernels (one capture each; ~40 s per capture)
for spec in "sins:ltv_fir_fft_kernel" "combsubfast:combsubfast_kernel"; do
  wl=${spec%%:*}; k=${spec##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$wl \
      python bench.py --workload $wl --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$wl.ncu-rep $k > gpurun_out/ncu_$wl.txt 2>&1
done
