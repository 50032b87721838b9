"""Drop-in `model/tsrn.py` for mjq11302010044/TATT: `from model import tsrn` (interfaces/base.py:20) keeps working with
ZERO edits to the callers, and `tsrn.TSRN` / `tsrn.TSRN_TL_TRANS` (interfaces/base.py:263,295) become the B200-native
modules of tatt_b200 (same constructor, forward signature, return structure and state_dict).

Install (in the reference checkout, with /path/to/tatt_b200's parent on PYTHONPATH):

    git mv model/tsrn.py model/tsrn_ref.py        # keep the original for the archs this path does not cover
    cp <tatt_b200 repo>/integration/model/tsrn.py model/tsrn.py

Everything else `model.tsrn` used to export (TSRN_C2F, SEM_TSRN, TSRN_TL, TSRN_TL_SFT, the building blocks ...) is
re-exported unchanged from `tsrn_ref` when that file exists.  There is no silent fallback for the two accelerated
classes: if the CUDA library is missing or the device is not a B200, constructing / calling them raises."""
try:                                            # the reference's original module, renamed by the maintainer
    from .tsrn_ref import *                      # noqa: F401,F403
    from . import tsrn_ref as _ref               # noqa: F401
except ImportError:                              # stand-alone use: only the accelerated classes exist
    _ref = None

from tatt_b200.tsrn import TSRN, TSRN_TL_TRANS   # noqa: E402,F401  (override the reference classes of the same name)

B200_NATIVE = ("TSRN", "TSRN_TL_TRANS")
