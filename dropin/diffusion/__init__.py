"""Drop-in replacement for the reference's `image_model/diffusion` package (`from diffusion import create_diffusion`,
inference.py:32, train_JPDVT.py:26).  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from jpdvt_mt_ntnu_b200.diffusion import create_diffusion, SpacedDiffusion, space_timesteps  # noqa: F401,E402
from jpdvt_mt_ntnu_b200.diffusion import gaussian_diffusion, respace  # noqa: F401,E402
from jpdvt_mt_ntnu_b200.diffusion import gaussian_diffusion as gd  # noqa: F401,E402

if _os.environ.get("JPDVT_DROPIN_VERBOSE"):
    import jpdvt_mt_ntnu_b200.diffusion as _d
    print(f"[jpdvt-dropin] diffusion -> {_d.__file__}", file=_sys.stderr, flush=True)
