"""Mirror of the reference's evaluation loop (``/root/reference/trainer/tester.py:11-62``) over the B200 path.

Same constructor and ``test()`` entry point: for every batch of the loader (``sample[0]`` = image tensor,
``sample[2]`` = list of sample_info dicts, ``data/dataset.py:84-87`` + the collate function) it runs model -> post-process
-> ``COCOMetrics.to_coco_format`` -> ``update_results``.  Differences: the three stages are timed with CUDA events
without a device synchronisation per stage (the reference's ``utils/timer.py`` synchronises around every block), the next
batch's host->device copy is issued on a side stream while the current batch computes, results are written with
``save_as_json``, and the AP accumulation (pycocotools' COCOeval, out of scope) only runs when pycocotools is importable.
The dataloader itself (cv2 decode / resize on CPU workers) stays the reference's code.
"""
import os

import torch

from .coco_format import COCOMetrics
from .visualizer import COCO_CAT_IDS      # the real COCO category ids (1..90 with gaps), data/dataset.py:42-49


class Tester:
    def __init__(self, model, postprocess, test_loader, checkpoint_dir, device, gt_file):
        self.model = model
        self.postprocess = postprocess
        self.test_loader = test_loader
        self.checkpoint_dir = checkpoint_dir
        self.device = torch.device(device)
        self.gt_file = gt_file
        dataset = getattr(test_loader, 'dataset', None)
        self.coco_metrics = COCOMetrics(gt_file=gt_file, cat2label=getattr(dataset, 'CAT2LABEL', None) or list(COCO_CAT_IDS),
                                        with_mask=getattr(dataset, 'with_mask', True), save_dir=checkpoint_dir)
        self.timings = {}

    def test(self):
        if self.device.type != 'cuda':
            raise RuntimeError('orienmask_b200.Tester runs on CUDA only; there is no CPU path')
        self.model.eval()
        copy_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        stages = {'Network Forward': [], 'Postprocess': [], 'Convert Format': []}
        images = 0

        def upload(sample):
            with torch.cuda.stream(copy_stream):
                img = sample[0].to(self.device, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            return img, sample[2], ready

        it = iter(self.test_loader)
        nxt = next(it, None)
        staged = upload(nxt) if nxt is not None else None
        with torch.no_grad():
            while staged is not None:
                image, batch_info, ready = staged
                nxt = next(it, None)
                staged = upload(nxt) if nxt is not None else None          # overlaps with this batch's compute
                cur.wait_event(ready)
                image.record_stream(cur)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
                predict = self.model(image)
                ev[1].record()
                detections = self.postprocess(predict)
                ev[2].record()
                coco_format_dets = self.coco_metrics.to_coco_format(batch_info, detections)
                ev[3].record()
                self.coco_metrics.update_results(coco_format_dets)
                images += int(image.shape[0])
                for name, a, b in zip(stages, ev[:-1], ev[1:]):
                    stages[name].append((a, b))
        torch.cuda.synchronize(self.device)
        self.timings = {name: sum(a.elapsed_time(b) for a, b in pairs) for name, pairs in stages.items()}
        self.timings['images'] = images
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        self.coco_metrics.save_as_json(os.path.join(self.checkpoint_dir, 'coco_format_results.json'))
        try:
            import pycocotools  # noqa: F401
            has_coco = bool(getattr(pycocotools, '__file__', None)) and self.gt_file is not None
        except ImportError:
            has_coco = False
        if has_coco:
            self.coco_metrics.coco_eval(per_cats=True)
        print('\n--------------------------------------------------------------------')
        print('Speed Statistics (%d images)' % images)
        for key in stages:
            ms = self.timings[key] / max(images, 1)
            print('%s: %.3fms (%.3ffps)' % (key, ms, 1000.0 / ms if ms > 0 else float('inf')))
        return self.timings
