"""Drop-in for the reference's ``model`` package (trainer/builder.py:12, infer.py:13 look classes up by name here)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from orienmask_b200.dropin._chain import chain_to_shadowed  # noqa: E402
# submodules that are not replaced here (eval.counter, eval.base, model.base, ...) resolve to the shadowed reference package
chain_to_shadowed(__name__, __path__, os.path.dirname(os.path.abspath(__file__)))

from orienmask_b200.model import OrienMaskYOLOFPNPlus, OrienMaskYOLO  # noqa: E402,F401
