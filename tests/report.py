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
            try:
                with open(PATH) as f:
                    data = json.load(f)
            except Exception:
                data = {}
        def plain(v):
            try:
                return float(v)
            except Exception:
                return str(v)
        data[test] = {k: plain(v) for k, v in metrics.items()}
        tmp = PATH + ".tmp"
        with open(tmp, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
        os.replace(tmp, PATH)
    except Exception:
        pass
