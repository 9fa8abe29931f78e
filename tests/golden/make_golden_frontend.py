"""Generate tests/golden/frontend_*.npz from the LIVE reference (build container only):

    python tests/golden/make_golden_frontend.py

Volume_Extractor is imported from the reference's ddsp/vocoder.py; cross_fade and the mask lines are executed from
main.py's OWN source text (main.py itself cannot be imported here: librosa / soundfile are not installed), compiled
function by function with `ast`, so the fixtures are outputs of the reference's code, not of the restatement."""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_function(py_file, name):
    """compile one top-level function of a reference source file without importing the file"""
    tree = ast.parse(open(os.path.join(ref_loader.REFERENCE_ROOT, py_file)).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), py_file, "exec"), ns)
    return ns[name]


def reference_mask_lines(volume, threshold_db):
    """main.py:211-213 executed from the file's own text (the three `mask = ...` statements)"""
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, "main.py")).read().splitlines()
    lines = [l.strip() for l in src if l.strip().startswith("mask = ") and "torch" not in l and "upsample" not in l]
    assert len(lines) == 3, lines

    class Cmd:
        threhold = threshold_db
    ns = {"np": np, "volume": volume, "cmd": Cmd}
    for l in lines:
        exec(l, ns)
    return ns["mask"]


def inputs(seed, T):
    g = np.random.default_rng(seed)
    t = np.arange(T) / 44100.0
    env = np.clip(np.sin(2 * np.pi * 0.7 * t + g.uniform(0, 6)), 0, None) ** 2          # bursts and silences
    return (env * 0.3 * np.sin(2 * np.pi * 220 * t) + 1e-4 * g.standard_normal(T)).astype(np.float32)


def main():
    V = ref_loader.load()[0]
    C = ref_loader.load()[1]
    cross_fade = reference_function("main.py", "cross_fade")
    for name, (seed, T, hop) in {"frontend_a": (1, 44100 + 77, 512), "frontend_b": (2, 20000, 441), "frontend_c": (3, 1500, 512)}.items():
        audio = inputs(seed, T)
        volume = V.Volume_Extractor(hop).extract(audio)
        mask = reference_mask_lines(volume, -40)
        m = torch.from_numpy(mask).float().unsqueeze(-1).unsqueeze(0)
        mask_up = C.upsample(m, hop).squeeze(-1).numpy()
        g = np.random.default_rng(seed + 10)
        a = g.standard_normal(5000).astype(np.float32)
        b = g.standard_normal(7000).astype(np.float32)
        idx = {"frontend_a": 3000, "frontend_b": 4999, "frontend_c": 0}[name]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), audio=audio, hop=np.int64(hop), volume=volume, mask=mask,
                            mask_up=mask_up.astype(np.float32), fade_a=a, fade_b=b, fade_idx=np.int64(idx),
                            fade_out=cross_fade(a, b, idx))
        print(name, volume.shape, volume.dtype, mask.sum(), mask_up.shape)


if __name__ == "__main__":
    main()
