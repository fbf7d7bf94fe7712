"""Host-side mirror of the reference's GPU pre-process, backed by one CUDA kernel (``om_preprocess``).

``FastCOCOTransform`` keeps the constructor and call signature of
``/root/reference/data/transform.py:444-510`` (``pipeline`` is the list of config dicts that
``trainer/builder.py:build_transform`` would turn into ``Resize`` / ``ShortEdgeResize`` / ``Normalize``
objects -- built transform objects with the same attribute names are accepted too) and ``pad`` mirrors
``/root/reference/infer.py:21-32``.  Input: ``[n, h, w, 3]`` float32 (as ``infer.py:148`` builds it) or
uint8 (what ``cv2.imread`` yields; same values, a quarter of the host->device bytes) on a CUDA device.
Output: contiguous fp32 NCHW.  The permute, the bilinear resize, the normalisation and (through
``transform_and_pad``) the padding are ONE pass over the output; there is no CPU path.
"""
import math

import torch

from . import _lib


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (list, tuple)) else (int(v), int(v))


def _get(step, name, default=None):
    return step.get(name, default) if isinstance(step, dict) else getattr(step, name, default)


def _kind(step):
    return step['type'] if isinstance(step, dict) else type(step).__name__


def pad_geometry(height, width, size_divisor=32):
    """infer.py:22-26 -> (pad_left, pad_right, pad_top, pad_down, new_height, new_width)."""
    new_height = int(math.ceil(height / size_divisor) * size_divisor)
    new_width = int(math.ceil(width / size_divisor) * size_divisor)
    pad_left, pad_top = (new_width - width) // 2, (new_height - height) // 2
    return [pad_left, new_width - width - pad_left, pad_top, new_height - height - pad_top, new_height, new_width]


def pad(image, size_divisor=32, pad_value=0):
    """infer.py:21-32 on an already transformed NCHW tensor (plain tensor plumbing; returns (image, pad_info))."""
    info = pad_geometry(image.shape[-2], image.shape[-1], size_divisor)
    if info[0] or info[1] or info[2] or info[3]:
        image = torch.nn.functional.pad(image, info[:4], value=pad_value)
    return image, info


class _Step:
    """Parameter holder with the reference's nested-class names, so ``trainer/builder.py:108-115`` (which builds every
    pipeline item as ``getattr(transform_class, item['type'])(**item)``) works on this class unchanged."""

    def __repr__(self):
        return '%s(%s)' % (type(self).__name__, ', '.join('%s=%r' % kv for kv in vars(self).items()))


class FastCOCOTransform:
    class Resize(_Step):                               # data/transform.py:463-474
        def __init__(self, size, interpolation='bilinear', align_corners=False):
            self.size, self.interpolation, self.align_corners = _pair(size), interpolation, align_corners

    class ShortEdgeResize(_Step):                      # data/transform.py:476-494
        def __init__(self, short_length, max_size, interpolation='bilinear', align_corners=False):
            self.short_length, self.max_size = short_length, max_size
            self.interpolation, self.align_corners = interpolation, align_corners

    class Normalize(_Step):                            # data/transform.py:496-507
        def __init__(self, mean, std):
            self.mean, self.std = mean, std

    def __init__(self, pipeline, use_cuda=True):
        if not use_cuda:
            raise RuntimeError('orienmask_b200.FastCOCOTransform runs on CUDA only; there is no CPU path')
        self.pipeline = list(pipeline)
        self.resize = None                 # ('fixed', (h, w)) | ('short', short_length, max_size)
        self.mean, self.std = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
        seen_norm = False
        for step in self.pipeline:
            kind = _kind(step)
            if kind in ('Resize', 'ShortEdgeResize'):
                if self.resize is not None or seen_norm:
                    raise NotImplementedError('one resize step, before Normalize, is supported')
                if _get(step, 'interpolation', 'bilinear') != 'bilinear' or _get(step, 'align_corners', False):
                    raise NotImplementedError('only bilinear interpolation with align_corners=False is implemented')
                if kind == 'Resize':
                    self.resize = ('fixed', _pair(_get(step, 'size')))
                else:
                    self.resize = ('short', _get(step, 'short_length'), _get(step, 'max_size'))
            elif kind == 'Normalize':
                if seen_norm:
                    raise NotImplementedError('a single Normalize step is supported')
                seen_norm = True
                self.mean = tuple(float(v) for v in _get(step, 'mean'))
                self.std = tuple(float(v) for v in _get(step, 'std'))
            else:
                raise NotImplementedError('unsupported FastCOCOTransform step %r' % kind)

    def output_size(self, h, w):
        if self.resize is None:
            return h, w
        if self.resize[0] == 'fixed':
            return self.resize[1]
        _, short_length, max_size = self.resize                      # data/transform.py:483-486
        scale = min(short_length / min(h, w), max_size / max(h, w))
        return int(h * scale + 0.5), int(w * scale + 0.5)

    def __call__(self, image, out=None):
        return self._run(image, None, 0.0, out)[0]

    def transform_and_pad(self, image, size_divisor=32, pad_value=0, out=None):
        """``pad(transform(image))`` of infer.py:149-150 in the same single pass; returns (image, pad_info).
        ``out``: optional preallocated fp32 [n, 3, H, W] result (a serving loop reuses its input buffer)."""
        return self._run(image, size_divisor, pad_value, out)

    def _run(self, image, size_divisor, pad_value, out=None):
        if not isinstance(image, torch.Tensor):
            raise RuntimeError('orienmask_b200.FastCOCOTransform needs a torch tensor, got %s' % type(image))
        if not image.is_cuda:
            # data/transform.py:456-457: the reference moves the image to its device (cuda:0) first; so does this (the work itself
            # has no CPU path -- without a CUDA device this raises)
            image = image.to('cuda:0')
        if image.dim() != 4 or image.shape[-1] != 3:
            raise ValueError('expected [n, h, w, 3], got %s' % (tuple(image.shape),))
        if image.dtype == torch.uint8:
            dtype = _lib.SRC_U8
        elif image.dtype == torch.float32:
            dtype = _lib.SRC_F32
        else:
            raise TypeError('expected a float32 or uint8 image tensor, got %s' % image.dtype)
        if image.stride(3) != 1 or image.stride(2) != 3 or image.stride(1) != 3 * image.shape[2]:
            image = image.contiguous()
        n, h, w, _ = image.shape
        rh, rw = self.output_size(h, w)
        info = pad_geometry(rh, rw, size_divisor) if size_divisor else [0, 0, 0, 0, rh, rw]
        cfg = _lib.PrepConfig()
        cfg.src_h, cfg.src_w, cfg.src_dtype = h, w, dtype
        cfg.resize_h, cfg.resize_w = rh, rw
        cfg.pad_top, cfg.pad_left, cfg.out_h, cfg.out_w = info[2], info[0], info[4], info[5]
        for c in range(3):
            cfg.mean[c], cfg.std[c] = self.mean[c], self.std[c]
        cfg.pad_value = float(pad_value)
        if out is None:
            out = torch.empty(n, 3, info[4], info[5], dtype=torch.float32, device=image.device)
        elif (tuple(out.shape) != (n, 3, info[4], info[5]) or out.dtype != torch.float32 or out.device != image.device
              or not out.is_contiguous()):
            raise ValueError('out must be a contiguous fp32 [%d, 3, %d, %d] tensor on %s' % (n, info[4], info[5], image.device))
        with torch.cuda.device(image.device):
            _lib.check(_lib.lib().om_preprocess(cfg, _lib.ptr(image), image.stride(0), n, _lib.ptr(out), _lib.stream_ptr()),
                       'om_preprocess')
        return out, info

    def __repr__(self):
        return '%s(resize=%s, mean=%s, std=%s)' % (type(self).__name__, self.resize, self.mean, self.std)
