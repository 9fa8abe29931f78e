"""Collect measured parity errors into gpurun_out/parity_report.json (read back in the build
container to calibrate tolerances and to quote in DESIGN.md / profiles)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "gpurun_out", "parity_report.json")


def record(test, **metrics):
    try:
        os.makedirs(os.path.dirname(PATH), exist_ok=True)
        data = {}
        if os.path.isfile(PATH):
            with open(PATH) as f:
                data = json.load(f)
        data[test] = {k: (float(v) if isinstance(v, (int, float)) else v) for k, v in metrics.items()}
        with open(PATH, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except Exception:
        pass
