"""``nms`` / ``batched_nms`` with the reference signatures (eval/function.py:55-103), on the CUDA NMS kernel.

The reference dispatches to a compiled ``nms_cuda.nms`` / ``nms_cpu.nms``; here both names go to the
C-ABI ``om_nms_ex`` (include/orienmask_b200.h).  ``semantics='cpu'`` (default) keeps the CPU variant's
rules -- IoU >= threshold suppresses, survivors are returned as ascending indices
(eval/src/nms_cpu.cpp:59,62): the variant the oracle is pinned to.  ``semantics='cuda'`` follows
eval/src/nms_kernel.cu (IoU > threshold, areas w*h, score-descending result), which is what the
reference itself runs on CUDA tensors; the default can be switched with ORIENMASK_B200_NMS=cuda.
CUDA tensors only; there is no CPU fallback.
"""
import os

import torch

from . import _lib


def default_semantics():
    return 'cuda' if os.environ.get('ORIENMASK_B200_NMS', 'cpu').lower() == 'cuda' else 'cpu'


def _mode(semantics):
    semantics = semantics or default_semantics()
    if semantics not in ('cpu', 'cuda'):
        raise ValueError("semantics must be 'cpu' or 'cuda'")
    return _lib.NMS_CUDA if semantics == 'cuda' else _lib.NMS_CPU


def _native_nms(dets, threshold, semantics=None):
    if not dets.is_cuda:
        raise RuntimeError('orienmask_b200 NMS needs a CUDA tensor; there is no CPU path')
    n = dets.size(0)
    if n > 1024:
        raise NotImplementedError('om_nms handles at most 1024 boxes per call (got %d)' % n)
    d = dets.detach().to(torch.float32).contiguous()
    keep = torch.empty(n, dtype=torch.int64, device=d.device)
    count = torch.empty(1, dtype=torch.int32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().om_nms_ex(_lib.ptr(d), n, float(threshold), _mode(semantics), _lib.ptr(keep), _lib.ptr(count),
                                        _lib.stream_ptr()), 'om_nms_ex')
    return keep[:int(count.item())]


def nms(dets, cats, threshold=0.5, semantics=None):
    """dets (n, 5): x, y, w, h, score; returns the kept dets, cats and their indices."""
    if dets.size(0) == 0:
        keep = dets.new_zeros(0, dtype=torch.long)
    else:
        keep = _native_nms(dets, threshold, semantics)
    return dets[keep], cats[keep], keep


def batched_nms(dets, cats, threshold=0.5, normalized=True, semantics=None):
    """Class-wise NMS: box centres are shifted by cls * (max_coordinate + 0.5) before plain NMS."""
    if dets.size(0) == 0:
        keep = dets.new_zeros(0, dtype=torch.long)
    else:
        max_coordinate = 1.5 if normalized else dets[:, :2].max() + dets[:, 2:4].max() / 2
        shifted = dets.clone()
        shifted[:, :2] += cats.float().view(-1, 1) * (max_coordinate + 0.5)
        keep = _native_nms(shifted, threshold, semantics)
    return dets[keep], cats[keep], keep
