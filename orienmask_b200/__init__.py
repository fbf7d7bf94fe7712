"""orienmask_b200 -- B200-native inference engine for OrienMask's hot path.

Public surface (mirrors the reference's names so its infer.py / test.py run unchanged through
``python -m orienmask_b200.dropin infer.py ...``):

    OrienMaskYOLOFPNPlus       model/orienmask_yolo_fpnplus.py   (forward pass on tcgen05 kernels)
    OrienMaskYOLO              model/orienmask_yolo.py           (the variant without skip convolutions)
    OrienMaskYOLOPostProcess   eval/orienmask_yolo_postprocess.py (decode + NMS + masks kernels)
    batched_nms, nms           eval/function.py
    FastCOCOTransform, pad     data/transform.py:444-510, infer.py:21-32 (pre-process kernel)
    COCOMetrics                eval/coco_eval.py:23-205 (mask crop/resize/RLE kernel; AP accumulation delegated to pycocotools)
    Tester                     trainer/tester.py:11-62 (evaluation loop over a dataloader)
    InferenceVisualizer        utils/visualizer.py:33-127 (mask area sort + alpha-blend kernels; boxes drawn by cv2)
"""
from .function import batched_nms, nms                      # noqa: F401
from .model import OrienMaskYOLOFPNPlus, OrienMaskYOLO       # noqa: F401
from .postprocess import OrienMaskYOLOPostProcess, PaddedDetections   # noqa: F401
from .transform import FastCOCOTransform, pad                 # noqa: F401
from .coco_format import COCOMetrics                          # noqa: F401
from .tester import Tester                                    # noqa: F401
from . import visualizer                                      # noqa: F401
from .visualizer import InferenceVisualizer                   # noqa: F401

__all__ = ['OrienMaskYOLOFPNPlus', 'OrienMaskYOLO', 'OrienMaskYOLOPostProcess', 'PaddedDetections', 'batched_nms', 'nms',
           'FastCOCOTransform', 'pad', 'COCOMetrics', 'Tester', 'InferenceVisualizer']
