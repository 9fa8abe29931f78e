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
# PCIe floor of the host-buffer (e2e) path: the step's 70.6 MB of pinned controls up, 56.4 MB of waveform down
python - <<'PY' | tee gpurun_out/pcie_floor.txt
import torch
up = torch.empty(70_636_440 // 4, dtype=torch.float32).pin_memory(); dn = torch.empty(56_426_496 // 4, dtype=torch.float32).pin_memory()
du, dd = torch.empty_like(up, device="cuda"), torch.empty_like(dn, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / n
h2d = t(lambda: du.copy_(up, non_blocking=True)); d2h = t(lambda: dn.copy_(dd, non_blocking=True))
def both():
    e = torch.cuda.Event(); e.record()
    with torch.cuda.stream(s1): s1.wait_event(e); du.copy_(up, non_blocking=True)
    with torch.cuda.stream(s2): s2.wait_event(e); dn.copy_(dd, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
print("H2D 70.6 MB: %.3f ms (%.1f GB/s)   D2H 56.4 MB: %.3f ms (%.1f GB/s)   both at once: %.3f ms" % (h2d, 70.636 / h2d, d2h, 56.426 / d2h, t(both)))
PY
