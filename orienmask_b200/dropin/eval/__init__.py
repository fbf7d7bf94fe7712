"""Drop-in for the hot-path names of the reference's ``eval`` package (trainer/builder.py:13-16)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from orienmask_b200.dropin._chain import chain_to_shadowed  # noqa: E402
# submodules that are not replaced here (eval.counter, eval.base, model.base, ...) resolve to the shadowed reference package
chain_to_shadowed(__name__, __path__, os.path.dirname(os.path.abspath(__file__)))

from orienmask_b200.postprocess import OrienMaskYOLOPostProcess  # noqa: E402,F401
from orienmask_b200.function import batched_nms, nms  # noqa: E402,F401
from . import function  # noqa: E402,F401
from .coco_eval import COCOMetrics  # noqa: E402,F401


def __getattr__(name):
    """``eval.EvalCounter`` (eval/__init__.py:2; running averages of the training loop) comes from the chained reference file."""
    if name == 'EvalCounter':
        from .counter import EvalCounter
        return EvalCounter
    raise AttributeError('drop-in eval package (orienmask_b200) has no attribute %r: only the inference path is replaced' % name)
