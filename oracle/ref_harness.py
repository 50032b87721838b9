"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (tatt_b200/).

Imports the live, unmodified reference (`/root/reference/model/tsrn.py`) in THIS
container so the oracle restatement (`oracle/tatt_oracle.py`) can be pinned against
it and golden fixtures can be generated (`tests/golden/make_golden.py`).

`/root/reference` does not exist on the GPU box: everything here is guarded by
`available()`; GPU-side tests use the committed fixtures + the oracle restatement.

The reference needs one stub to import (`model/tsrn.py:9` does `from IPython import
embed`; IPython is absent) -- injected here, reference files untouched.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("TATT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "tsrn.py"))


_mod = None


def load():
    """Return the reference `model.tsrn` module (cached)."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True  # the reference dir is read-only
    if "IPython" not in sys.modules:
        stub = types.ModuleType("IPython")
        stub.embed = lambda *a, **k: None
        sys.modules["IPython"] = stub
    import warnings
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib
        _mod = importlib.import_module("model.tsrn")
    return _mod


def zero_dropout(module):
    """Determinism recipe of SURVEY 8c: every nn.Dropout.p = 0 and every
    nn.MultiheadAttention.dropout = 0 -> bit-identical train-mode repeats."""
    import torch.nn as nn
    for m in module.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return module


def perturb_(module, seed=7):
    """Make random-init parity meaningful: randomise BN running stats and
    affine, un-zero stn_fc2.weight (zero-init hides all STN gradients)."""
    import torch
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in module.named_modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            if name.endswith("stn_fc2"):
                m.weight.copy_(0.02 * torch.randn(m.weight.shape, generator=g))
    return module
