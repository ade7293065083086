"""Import the UNMODIFIED reference from /root/reference (this container only).

TEST INFRASTRUCTURE. Used by `oracle/make_golden.py` and by the `-m "not gpu"`
tests that pin the oracle restatement against the real reference when the
read-only mount is present.  Nothing on the GPU box may import this module:
/root/reference does not exist there.

Three shims, none of which edits a reference file:
  * `matplotlib` / `matplotlib.pyplot` stub modules (src/utils/utils.py:11
    imports pyplot at module top; it is not installed here);
  * `np.int = int; np.float = float` (src/utils/utils.py:292-293 use the
    aliases NumPy removed in 1.24);
  * the DR-SPAAM model is imported from `src.depracted.model.dr_spaam`
    directly, because `src/depracted/model/__init__.py` is empty
    (bin/eval_dr_spaam.py:18 cannot work as written).
"""
import os
import sys
import types

_INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _pick_root():
    """The read-only mount in the build container; on the GPU box the unmodified copy that baseline/install_reference.py
    placed under the git-ignored baseline/_ref/ (used by `bench.py --impl reference` only)."""
    for cand in (os.environ.get("POF_REFERENCE_ROOT"), "/root/reference", _INSTALLED):
        if cand and os.path.isfile(os.path.join(cand, "src", "utils", "utils.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "utils", "utils.py"))


_cache = {}


def load():
    """Return (utils_module, dr_spaam_module) of the reference."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np

    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float

    # The reference is a top-level package called `src`; this repo also ships a
    # `src` compatibility package, so the reference gets loaded under its own
    # sys.path entry with any previously imported `src*` modules parked aside.
    parked = {k: sys.modules.pop(k) for k in list(sys.modules)
              if k == "src" or k.startswith("src.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib

        ref_utils = importlib.import_module("src.utils.utils")
        ref_model = importlib.import_module("src.depracted.model.dr_spaam")
        _cache["prototype"] = importlib.import_module("src.depracted.model.prototype")
        try:
            _cache["dataset_dr_spaam"] = importlib.import_module("src.utils.dataset_dr_spaam")
        except Exception as e:      # noqa: BLE001  (only the dataset test needs it; it reports the reason)
            _cache["dataset_dr_spaam"] = e
    finally:
        sys.path.remove(REFERENCE_ROOT)
        ref_mods = {k: sys.modules.pop(k) for k in list(sys.modules)
                    if k == "src" or k.startswith("src.")}
        sys.modules.update(parked)
    _cache["mods"] = (ref_utils, ref_model)
    _cache["all"] = ref_mods
    return _cache["mods"]


def load_prototype():
    """The reference's `src.depracted.model.prototype` module (Prototype, flow_loss)."""
    load()
    return _cache["prototype"]


def load_dataset_module():
    """The reference's `src.utils.dataset_dr_spaam` module (DROWDataset2 and the loaders)."""
    load()
    mod = _cache["dataset_dr_spaam"]
    if isinstance(mod, Exception):
        raise mod
    return mod


class cpu_cuda_noop:
    """`Prototype._fusion` allocates a scratch tensor with `.cuda()` that it never uses
    (prototype.py:121); on the GPU-less build container that one call is made a no-op while the
    reference runs.  Nothing else about the reference is altered."""

    def __enter__(self):
        import torch

        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch

        torch.Tensor.cuda = self._orig
        return False
