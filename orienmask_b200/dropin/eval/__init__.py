"""Drop-in for the hot-path names of the reference's ``eval`` package (trainer/builder.py:13-16)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from orienmask_b200.postprocess import OrienMaskYOLOPostProcess  # noqa: E402,F401
from orienmask_b200.function import batched_nms, nms  # noqa: E402,F401
from . import function  # noqa: E402,F401
