"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's GPU pre-process.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.

Follows ``/root/reference/data/transform.py:444-510`` (``FastCOCOTransform``: ``permute(0,3,1,2)`` :459,
``Resize`` :463-474 / ``ShortEdgeResize`` :476-494 = ``F.interpolate(mode='bilinear',
align_corners=False)``, ``Normalize`` :496-507 = ``sub_(mean).div_(std)``) and
``/root/reference/infer.py:21-32`` (``pad``).  The interpolation arithmetic lives in PyTorch (ATen
``upsample_bilinear2d``, un-pinned ``torch`` in requirements.txt) and is not a single rounding sequence: the
CPU kernel of torch 2.11 contracts ``a*b + c*d`` differently (or not at all) depending on the loop instance a
shape lands in, and infer.py runs the CUDA kernel (``use_cuda=True``).  This file restates the CUDA kernel's
expression with nvcc's default contraction (identical to what torch 2.11 CPU computes for the 544x544
up-scales); ``tests/test_prep.py`` pins it against ``torch.nn.functional.interpolate`` (bit-exact at 544,
<= 4e-3 on 0..255 data elsewhere) and against ``tests/golden/prep_small.npz``, the output of the reference's own
``FastCOCOTransform`` + ``pad`` (``tests/golden/make_golden_prep.py``).

Arithmetic (fp32, each op single-rounded): scale = in/out; src = max(fma(scale, d + 0.5, -0.5), 0);
i0 = (int)src; i1 = min(i0 + 1, in - 1); l1 = src - i0; l0 = 1 - l1;
row = fma(l0x, v[i0], l1x * v[i1]); out = fma(l0y, row0, l1y * row1); then (out - mean) / std.
"""
import math

import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def _lerp_index(n_out, n_in):
    scale = f32(n_in) / f32(n_out)
    d = np.arange(n_out, dtype=f32)
    src = np.maximum(_fma(np.full(n_out, scale, f32), d + f32(0.5), np.full(n_out, -0.5, f32)), f32(0))
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    l0 = (f32(1) - l1).astype(f32)
    return i0, i1, l0, l1


def bilinear_resize(x, out_h, out_w):
    """[..., h, w] fp32 -> [..., out_h, out_w] (data/transform.py:470-472)."""
    x = np.asarray(x, dtype=f32)
    h, w = x.shape[-2:]
    y0, y1, ly0, ly1 = _lerp_index(out_h, h)
    x0, x1, lx0, lx1 = _lerp_index(out_w, w)
    rows0, rows1 = x[..., y0, :], x[..., y1, :]
    shp = rows0[..., x0].shape
    LX0, LX1 = np.broadcast_to(lx0, shp), np.broadcast_to(lx1, shp)
    r0 = _fma(LX0, rows0[..., x0], LX1 * rows0[..., x1])
    r1 = _fma(LX0, rows1[..., x0], LX1 * rows1[..., x1])
    LY0, LY1 = np.broadcast_to(ly0[:, None], shp), np.broadcast_to(ly1[:, None], shp)
    return _fma(LY0, r0, LY1 * r1)


def short_edge_size(h, w, short_length, max_size):
    """data/transform.py:483-486."""
    scale = min(short_length / min(h, w), max_size / max(h, w))
    return int(h * scale + 0.5), int(w * scale + 0.5)


def fast_transform_oracle(image_nhwc, size=None, mean=(0, 0, 0), std=(255, 255, 255)):
    """FastCOCOTransform([Resize(size), Normalize(mean, std)]) on [n, h, w, 3] -> fp32 [n, 3, H, W]."""
    x = np.ascontiguousarray(np.asarray(image_nhwc).astype(f32).transpose(0, 3, 1, 2))      # :459
    if size is not None:
        x = bilinear_resize(x, int(size[0]), int(size[1]))
    m = np.asarray(mean, f32)[None, :, None, None]
    s = np.asarray(std, f32)[None, :, None, None]
    return ((x - m).astype(f32) / s).astype(f32)                                               # :505


def pad_oracle(image, size_divisor=32, pad_value=0):
    """infer.py:21-32 -> (padded image, [left, right, top, down, new_h, new_w])."""
    h, w = image.shape[-2:]
    nh = int(math.ceil(h / size_divisor) * size_divisor)
    nw = int(math.ceil(w / size_divisor) * size_divisor)
    left, top = (nw - w) // 2, (nh - h) // 2
    right, down = nw - w - left, nh - h - top
    out = np.full(image.shape[:-2] + (nh, nw), f32(pad_value), dtype=f32)
    out[..., top:top + h, left:left + w] = image
    return out, [left, right, top, down, nh, nw]
