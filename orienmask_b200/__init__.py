"""orienmask_b200 -- B200-native inference engine for OrienMask's hot path.

Public surface (mirrors the reference's names so its infer.py / test.py run unchanged when
``orienmask_b200/dropin`` is first on PYTHONPATH):

    OrienMaskYOLOFPNPlus       model/orienmask_yolo_fpnplus.py   (forward pass on tcgen05 kernels)
    OrienMaskYOLOPostProcess   eval/orienmask_yolo_postprocess.py (decode + NMS + masks kernels)
    batched_nms, nms           eval/function.py
"""
from .function import batched_nms, nms                      # noqa: F401
from .model import OrienMaskYOLOFPNPlus                      # noqa: F401
from .postprocess import OrienMaskYOLOPostProcess, PaddedDetections   # noqa: F401

__all__ = ['OrienMaskYOLOFPNPlus', 'OrienMaskYOLOPostProcess', 'PaddedDetections', 'batched_nms', 'nms']
