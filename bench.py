#!/usr/bin/env python
"""Benchmark of the DDSP synthesis hot path (BASELINE.json metric: audio Msamples/s at 44.1 kHz,
128 harmonics; % of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload sins|...]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE).  A "step" is one pass of the
Sins synthesis path (frame phase scan -> sinusoid bank -> impulse responses -> two time-varying
FIRs + mix, in-kernel Philox noise) over one batch of B=32 utterances x 10 s per GPU
(BASELINE configs[1]; configs[3] at N=8 = 256 utterances); at N>1 the step also gathers every
rank's `signal` to rank 0 over NCCL (the path's only collective).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, P = 44100, 512
WORKLOADS = {
    # name: (kind, batch per GPU, seconds, params, algorithmic bytes per output sample)
    "sins": dict(kind="sins", B=32, sec=10, H=128, Ma=256, Mn=256,
                 label="Sins forward DSP, B=32 x 10 s x 128 harmonics, 44.1 kHz, n_mag 256/256 (BASELINE configs[1])"),
    "sins_cfg1": dict(kind="sins", B=1, sec=2, H=64, Ma=256, Mn=256,
                      label="Sins forward DSP, B=1 x 2 s x 64 harmonics (BASELINE configs[0])"),
}


def algorithmic_bytes(w, nF):
    """bytes the path must move per launch of the whole path: controls + f0 read once, every
    returned tensor written once."""
    B, T = w["B"], nF * P
    if w["kind"] == "sins":
        c = w["H"] + w["Ma"] + w["Mn"]
        return 4 * B * nF * (1 + c) + 4 * B * T * 3
    raise ValueError(w["kind"])


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------
def cpu_reference_run(w, batch, reps, threads=None):
    """Time the reference's CPU algorithm (oracle port on the same ATen CPU operators) on
    `batch` utterances of the workload.  Returns (Msamples/s, seconds per rep, cores)."""
    import torch
    from ddsp_svc_b200 import synthetic as syn
    from oracle import torch_port as tp
    cores = threads or best_thread_count(w)
    torch.set_num_threads(cores)
    nF = syn.n_frames_for(w["sec"], SR, P)
    sm = syn.sins_split_map(w["H"], w["Ma"], w["Mn"])
    f0 = syn.make_f0(batch, nF, SR, P)
    _, ctrls = syn.make_ctrl(batch, nF, sm)
    best = None
    with torch.no_grad():
        tp.sins_forward(f0[:1], {k: v[:1] for k, v in ctrls.items()}, SR, P)  # warm-up (small)
        for _ in range(reps):
            t0 = time.perf_counter()
            tp.sins_forward(f0, ctrls, SR, P)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return batch * nF * P / best / 1e6, best, cores


_best_threads = {}


def best_thread_count(w):
    """The reference runs PyTorch with its default intra-op pool (all cores).  On many-core hosts
    that is slower than a smaller pool for these medium-sized ops, so give the CPU arm its best
    setting: try a few pool sizes on one utterance and keep the fastest."""
    import torch
    from ddsp_svc_b200 import synthetic as syn
    from oracle import torch_port as tp
    key = w["label"]
    if key in _best_threads:
        return _best_threads[key]
    ncpu = os.cpu_count() or 1
    nF = syn.n_frames_for(min(w["sec"], 2), SR, P)
    sm = syn.sins_split_map(w["H"], w["Ma"], w["Mn"])
    f0 = syn.make_f0(2, nF, SR, P)
    _, ctrls = syn.make_ctrl(2, nF, sm)
    best, best_t = ncpu, None
    for n in sorted({c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu}):
        torch.set_num_threads(n)
        with torch.no_grad():
            tp.sins_forward(f0, ctrls, SR, P)
            t0 = time.perf_counter()
            tp.sins_forward(f0, ctrls, SR, P)
            dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    _best_threads[key] = best
    return best


def run_reference_arm(args, w):
    """--impl reference: the reference's own CPU implementation of the path on this box's host
    cores.  /root/reference (Python) cannot travel to the GPU box, so this times the oracle port,
    which is bit-identical to it on CPU (tests/test_oracle_vs_reference.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ddsp_svc_b200 import synthetic as syn
    nF = syn.n_frames_for(w["sec"], SR, P)
    sample_b = min(w["B"], 4)
    vals = []
    for _ in range(args.warmup):
        cpu_reference_run(w, sample_b, 1)
    t_total = 0.0
    for _ in range(args.steps):
        v, dt, cores = cpu_reference_run(w, sample_b, 1)
        vals.append(v); t_total += dt
    value = sample_b * nF * P * args.steps / t_total / 1e6
    sample = "%d of %d utterances x %d s per step (bounded sample of the same workload)" % (sample_b, w["B"], w["sec"])
    line = {"impl": "reference", "metric": "audio Msamples/s (44.1 kHz, %d harmonics)" % w["H"], "value": value,
            "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "sample": sample, "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sins", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="also print per-kernel times to stderr")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, w)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from ddsp_svc_b200 import FixedControls, Sins, ops, synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = w["B"]
    nF = syn.n_frames_for(w["sec"], SR, P)
    T = nF * P
    sm = syn.sins_split_map(w["H"], w["Ma"], w["Mn"])
    # per-rank shard of the global batch: utterances [rank*B, (rank+1)*B)
    f0_h = syn.make_f0(B, nF, SR, P, seed=1234 + rank).pin_memory()
    dense_h, _ = syn.make_ctrl(B, nF, sm, seed=7 + rank)
    dense_h = dense_h.pin_memory()
    f0_d = f0_h.to(dev)
    dense_d = dense_h.to(dev)
    ctrl_d = syn.split_views(dense_d, sm)
    hidden = torch.zeros(B, nF, 256, device=dev)
    fixed = FixedControls(ctrl_d, hidden)
    model = Sins(SR, P, w["H"], w["Ma"], w["Mn"], unit2ctrl=fixed).to(dev)
    gathered = torch.empty(world * B, T, device=dev) if (world > 1 and rank == 0) else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out_h = torch.empty(B, T, dtype=torch.float32).pin_memory()

    def step():
        sig, _, (harm, nz) = model(None, f0_d, None, utterance_offset=rank * B)
        if world > 1:
            dist.gather(sig, list(gathered.split(B)) if rank == 0 else None, dst=0)
        return sig

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        sync_all()
        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        n0 = ops.launches()
        sync_all()
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()                      # evict L2 between steps (untimed)
            a.record()
            step()
            b.record()
        sync_all()
        wall = time.perf_counter() - wall0
        launches = ops.launches() - n0
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)

        # ---- end to end through the public module API with HOST buffers ----
        e2e_steps = max(3, min(args.steps, 10))
        h2d = f0_h.numel() * 4 + dense_h.numel() * 4
        d2h = out_h.numel() * 4
        ee = []
        for i in range(e2e_steps + 1):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            f0_x = f0_h.to(dev, non_blocking=True)
            dn_x = dense_h.to(dev, non_blocking=True)
            fixed.ctrls = syn.split_views(dn_x, sm)
            sig, _, _ = model(None, f0_x, None, utterance_offset=rank * B)
            out_h.copy_(sig, non_blocking=True)
            b.record()
            b.synchronize()
            if i > 0:
                ee.append(a.elapsed_time(b))
        fixed.ctrls = ctrl_d
        e2e_ms = sum(ee) / len(ee)
        clk = clocks.stop() if rank == 0 else None

        # ---- per-kernel durations (CUDA events on the launching stream), for the roofline ----
        kt = {}
        reps = max(5, min(args.steps, 20))
        fp, _ = ops.phase_scan(f0_d, P, SR)
        sinus = ops.sins_bank(f0_d, fp, ctrl_d["amplitudes"], P, SR)
        ir_a = ops.ir_build(ctrl_d["group_delay"], ops.IR_ALLPASS, SR)
        ir_n = ops.ir_build(ctrl_d["noise_magnitude"], ops.IR_MAG_HANN, SR)
        L = ops._lib.lib()
        sig_b, har_b, nz_b = (torch.empty(B, T, device=dev) for _ in range(3))
        st = torch.cuda.current_stream().cuda_stream

        def timed(name, fn):
            ts = []
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); b.synchronize()
                ts.append(a.elapsed_time(b))
            kt[name] = sum(ts) / len(ts)

        timed("phase_scan", lambda: ops.phase_scan(f0_d, P, SR))
        timed("sins_bank", lambda: ops.sins_bank(f0_d, fp, ctrl_d["amplitudes"], P, SR))
        timed("ir_build_allpass", lambda: ops.ir_build(ctrl_d["group_delay"], ops.IR_ALLPASS, SR))
        timed("ir_build_noise", lambda: ops.ir_build(ctrl_d["noise_magnitude"], ops.IR_MAG_HANN, SR))
        timed("ltv_fir_x2_mix", lambda: L.b2d_ltv_fir(sinus.data_ptr(), ir_a.data_ptr(), 510, har_b.data_ptr(), 0,
                                                      ir_n.data_ptr(), 510, nz_b.data_ptr(), sig_b.data_ptr(), 1, 0,
                                                      B, nF, P, st))

    # ---- reduce over ranks: max device time ----
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t[0].item(), t[1].item()
    ms_per_step = dev_ms / args.steps
    samples_step = world * B * T
    value = samples_step / (ms_per_step * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        dom = max(kt, key=kt.get)
        alg_bytes = algorithmic_bytes(w, nF)
        achieved = alg_bytes / (kt[dom] * 1e-3) / 1e9
        line = {
            "metric": "audio Msamples/s (44.1 kHz, %d harmonics)" % w["H"],
            "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "per_gpu_batch": B, "global_batch": world * B, "n_frames": nF,
                       "samples_per_utterance": T, "noise": "in-kernel Philox4x32-10",
                       "outputs": "signal+harmonic+noise", "parallelism": "batch-sharded x%d%s" % (
                           world, ", NCCL gather of signal to rank 0 inside the step" if world > 1 else ""),
                       "l2": "flushed between steps (256 MiB memset, untimed); per-step CUDA events summed",
                       "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kt,
                         "whole_path_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak if world == 1 else None,
                         "note": "Sins is FP32/SFU-issue bound (2040 FMA + 70 MUFU per sample against 17 B); "
                                 "see DESIGN.md for the pipe-utilisation view"},
            "e2e": {"value": samples_step / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "pinned host f0+controls -> H2D -> Sins.forward (public module API) -> D2H of signal"},
            "gpu_launches": launches,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            sample_b = 4
            v, dt, cores = cpu_reference_run(w, sample_b, 3)
            line["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                    "sample": "%d of %d utterances x %d s, best of 3 (%.2f s per pass)" % (
                                        sample_b, B, w["sec"], dt)}
        if args.breakdown:
            print(json.dumps(kt), file=sys.stderr)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
