"""Drop-in replacement for the reference's `image_model/models.py`.

Put this directory FIRST on PYTHONPATH (or copy the two shims next to the reference scripts) and the reference's
`from models import DiT_models, get_2d_sincos_pos_embed` (inference.py:31, train_JPDVT.py:24) resolves to the B200
implementation.  See INTEGRATION.md.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from jpdvt_mt_ntnu_b200.models import *  # noqa: F401,F403,E402
from jpdvt_mt_ntnu_b200.models import DiT, DiT_models, get_2d_sincos_pos_embed  # noqa: F401,E402

if _os.environ.get("JPDVT_DROPIN_VERBOSE"):
    import jpdvt_mt_ntnu_b200.models as _m
    print(f"[jpdvt-dropin] models -> {_m.__file__}", file=_sys.stderr, flush=True)
