"""Drop-in for ``eval.coco_eval`` of the reference (infer.py:17, trainer/tester.py): COCOMetrics whose
``to_coco_format`` runs the crop / resize / RLE kernel instead of per-instance D2H + pycocotools."""
from orienmask_b200.coco_format import COCOMetrics  # noqa: F401
