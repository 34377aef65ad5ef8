"""Import shim for the real reference (TEST INFRASTRUCTURE ONLY).

Loads ``/root/reference/networks/MSTr.py`` without touching the reference tree.
It exists only in the authoring container: it validates ``oracle/mstr_oracle.py``
and generates ``tests/golden/`` fixtures (``oracle/make_golden.py``).  Nothing in
the product path, ``bench.py`` or the ``-m gpu`` tests may import this file at run
time on the GPU box (``/root/reference`` does not exist there).

Two shims:
  * ``torchinfo`` is not installed, ``MSTr.py:13`` imports it -> stub module.
  * ``silu_sigmoid.forward`` (``MSTr.py:1275-1277``) hard-codes ``.cuda()`` ->
    replaced by a device-agnostic equivalent ``min(SiLU(x+3)/6, 1)``.
"""
import os
import sys
import types
import importlib

REF_ROOT = os.environ.get("TRANSCEPTION_REF", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "networks", "MSTr.py"))


def load_reference():
    """Return the reference ``networks.MSTr`` module object (shimmed)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if "torchinfo" not in sys.modules:
        stub = types.ModuleType("torchinfo")
        stub.summary = lambda *a, **k: None
        sys.modules["torchinfo"] = stub
    name = "_tcx_ref_networks"
    if name + ".MSTr" in sys.modules:
        return sys.modules[name + ".MSTr"]
    pkg = types.ModuleType(name)
    pkg.__path__ = [os.path.join(REF_ROOT, "networks")]
    sys.modules[name] = pkg
    mod = importlib.import_module(name + ".MSTr")
    import torch

    def _silu_sigmoid_forward(self, x):
        x = self.silu(x + 3) / 6
        return torch.minimum(x, torch.ones_like(x))

    mod.silu_sigmoid.forward = _silu_sigmoid_forward
    return mod
