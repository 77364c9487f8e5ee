"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference motion module.

Used in the build container (where /root/reference exists) to
  * pin oracle/motion_oracle.py against the real reference (tests/test_oracle_pin.py), and
  * generate the committed golden fixtures (oracle/gen_golden.py -> tests/golden/).
It never travels to the GPU box (the reference tree is absent there) and nothing in the product
package (neurons_b200/) may import it.

The reference file animatediff/models/motion_module.py:10-12 imports three names from
`diffusers` (0.11.1 API), which is not installed in this image:
    diffusers.utils.BaseOutput, diffusers.utils.import_utils.is_xformers_available,
    diffusers.models.attention.{CrossAttention, FeedForward}
The reference tree carries a verbatim in-tree copy of the last two in
animatediff/models/motion_module_new.py:119-339,429-534 (dead code upstream, nothing imports it).
We register stub `diffusers` modules in sys.modules, load motion_module_new.py by path to obtain
CrossAttention/FeedForward, expose them as diffusers.models.attention.*, then load the live
motion_module.py by path.  No reference source is copied or modified.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from collections import OrderedDict

_REF_ENV = "NEURONS_REF"
_DEFAULT_REF = "/root/reference"
_cached = None


def reference_root() -> str | None:
    """$NEURONS_REF, /root/reference, or <repo>/baseline/_ref (the base contract's install target; empty here: the reference is pure
    Python without a setup.py / pyproject, so there is nothing to pip-install -- see DESIGN.md section 7)."""
    repo_ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    for cand in (os.environ.get(_REF_ENV), _DEFAULT_REF, repo_ref):
        if cand and os.path.isfile(os.path.join(cand, "animatediff", "models", "motion_module.py")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def _load_by_path(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_motion_module():
    """Return the reference's `motion_module` python module (get_motion_module, VanillaTemporalModule ...)."""
    global _cached
    if _cached is not None:
        return _cached
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set NEURONS_REF or mount /root/reference)")

    if "diffusers" not in sys.modules:
        class BaseOutput(OrderedDict):
            """dataclass-friendly stand-in for diffusers.utils.BaseOutput (only subclassed, never used)."""

        d = types.ModuleType("diffusers")
        du = types.ModuleType("diffusers.utils")
        dui = types.ModuleType("diffusers.utils.import_utils")
        dm = types.ModuleType("diffusers.models")
        dma = types.ModuleType("diffusers.models.attention")
        du.BaseOutput = BaseOutput
        dui.is_xformers_available = lambda: False
        du.import_utils = dui
        d.utils = du
        d.models = dm
        dm.attention = dma
        for m in (d, du, dui, dm, dma):
            sys.modules[m.__name__] = m
    base = os.path.join(root, "animatediff", "models")
    mm_new = _load_by_path("_neurons_ref_motion_module_new", os.path.join(base, "motion_module_new.py"))
    dma = sys.modules["diffusers.models.attention"]
    dma.CrossAttention = mm_new.CrossAttention
    dma.FeedForward = mm_new.FeedForward
    _cached = _load_by_path("_neurons_ref_motion_module", os.path.join(base, "motion_module.py"))
    return _cached
