"""``nms`` / ``batched_nms`` with the reference signatures (eval/function.py:55-103), on the CUDA NMS kernel.

The reference dispatches to a compiled ``nms_cuda.nms`` / ``nms_cpu.nms``; here both names go to the
C-ABI ``om_nms`` (include/orienmask_b200.h), which keeps the CPU variant's semantics -- IoU >=
threshold suppresses, survivors are returned as ascending indices (eval/src/nms_cpu.cpp:59,62).
CUDA tensors only; there is no CPU fallback.
"""
import torch

from . import _lib


def _native_nms(dets, threshold):
    if not dets.is_cuda:
        raise RuntimeError('orienmask_b200 NMS needs a CUDA tensor; there is no CPU path')
    n = dets.size(0)
    if n > 1024:
        raise NotImplementedError('om_nms handles at most 1024 boxes per call (got %d)' % n)
    d = dets.detach().to(torch.float32).contiguous()
    keep = torch.empty(n, dtype=torch.int64, device=d.device)
    count = torch.empty(1, dtype=torch.int32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().om_nms(_lib.ptr(d), n, float(threshold), _lib.ptr(keep), _lib.ptr(count),
                                     _lib.stream_ptr()), 'om_nms')
    return keep[:int(count.item())]


def nms(dets, cats, threshold=0.5):
    """dets (n, 5): x, y, w, h, score; returns the kept dets, cats and their indices."""
    if dets.size(0) == 0:
        keep = dets.new_zeros(0, dtype=torch.long)
    else:
        keep = _native_nms(dets, threshold)
    return dets[keep], cats[keep], keep


def batched_nms(dets, cats, threshold=0.5, normalized=True):
    """Class-wise NMS: box centres are shifted by cls * (max_coordinate + 0.5) before plain NMS."""
    if dets.size(0) == 0:
        keep = dets.new_zeros(0, dtype=torch.long)
    else:
        max_coordinate = 1.5 if normalized else dets[:, :2].max() + dets[:, 2:4].max() / 2
        shifted = dets.clone()
        shifted[:, :2] += cats.float().view(-1, 1) * (max_coordinate + 0.5)
        keep = _native_nms(shifted, threshold)
    return dets[keep], cats[keep], keep
