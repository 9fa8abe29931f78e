#!/usr/bin/env python
"""Stage the UNMODIFIED reference (yxlllc/DDSP-SVC, a pure-Python program) under baseline/_ref/ so that it travels to the
GPU box with the repo snapshot (baseline/_ref/ is git-ignored, not gpurun-ignored; /root/reference itself exists only in
the build container).  Only Python sources and YAML configs are staged (about 1 MB); nothing is edited.

It is used there as a CHECKER and a BASELINE only:
  * tests/test_gpu_acceptance.py runs the reference's own classes (real Unit2Control, real torch.split views, spk_id /
    spk_mix_dict) on the GPU next to the patch_reference() drop-in and compares the waveforms;
  * bench.py times the reference's Sins.forward eagerly on the same GPU (`eager_gpu_baseline.kind: "reference"`).
Nothing in ddsp_svc_b200/ imports it.

    python tools/stage_reference.py [--src /root/reference] [--dst baseline/_ref]
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage(src="/root/reference", dst=None, quiet=False):
    dst = dst or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(src, "ddsp", "vocoder.py")):
        return None
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__", "exp", "data", "pretrain")]
        rel = os.path.relpath(base, src)
        for f in files:
            if not f.endswith((".py", ".yaml")):
                continue
            out_dir = os.path.join(dst, rel) if rel != "." else dst
            os.makedirs(out_dir, exist_ok=True)
            s, d = os.path.join(base, f), os.path.join(out_dir, f)
            if not os.path.isfile(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
                shutil.copy2(s, d)
            n += 1
    if not quiet:
        print("staged %d reference files under %s" % (n, dst))
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=None)
    a = ap.parse_args()
    sys.exit(0 if stage(a.src, a.dst) else 1)
