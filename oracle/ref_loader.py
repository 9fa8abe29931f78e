"""Import the live reference (yxlllc/DDSP-SVC) for oracle validation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is a Python /
PyTorch program; its hot path (ddsp/vocoder.py, ddsp/core.py,
nsf_hifigan/models.py) imports several third-party packages that are not
installed here and are not used by the path.  They are replaced by empty stub
modules before import.  /root/reference exists only in the build container; on the
GPU box the staged copy under baseline/_ref/ (tools/stage_reference.py) is used when
present, so callers must check ``available()`` first.
"""
import os
import sys
import types

def _find_root():
    """The live checkout in the build container, else the staged copy that travels to the GPU box
    (baseline/_ref/, written by tools/stage_reference.py; git-ignored)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("DDSP_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "ddsp", "vocoder.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_STUBS = ("pyworld", "parselmouth", "torchcrepe", "resampy", "fairseq", "gin",
          "local_attention")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ddsp", "vocoder.py"))


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["fairseq"].checkpoint_utils = types.SimpleNamespace()
    sys.modules["local_attention"].LocalAttention = object
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.use = lambda *a, **k: None
            sys.modules["matplotlib"] = mpl
            for sub in ("pylab", "pyplot"):
                sys.modules["matplotlib." + sub] = types.ModuleType("matplotlib." + sub)


_cache = {}


def load():
    """Return (ddsp.vocoder module, ddsp.core module, nsf_hifigan.models.SineGen)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("live reference not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ddsp.vocoder as V
    import ddsp.core as C
    from nsf_hifigan.models import SineGen
    _cache["mods"] = (V, C, SineGen)
    return _cache["mods"]


def fixed_ctrl_module(ctrls, hidden):
    """A torch module standing in for Unit2Control: returns fixed raw controls.

    This is the DSP/NN seam of the reference (ddsp/vocoder.py:578, :664, :832).
    """
    import torch

    class FixedCtrl(torch.nn.Module):
        def forward(self, *a, **k):
            return ctrls, hidden

    return FixedCtrl()
