"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference visualiser's mask blend.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.

Follows ``/root/reference/utils/visualizer.py``: ``_recover_shape_segm`` :122-127 (crop ``pad_info``, bilinear resize, NOT
rounded), the area sort :69 and ``plot_all_mask`` :95-100.  The resize is ``oracle/prep_oracle.bilinear_resize`` (see there
for why the interpolation has no single bit-exact definition); the k >= 1 terms are added with ``np.sum`` where the
reference uses ``torch.sum`` (order unspecified), so parity is to fp32 rounding: pinned by ``tests/golden/blend_small.npz``
(made by the reference's own methods, tests/golden/make_golden_blend.py) to 1e-3 on 0..255 data.
"""
import numpy as np

from .prep_oracle import bilinear_resize

f32 = np.float32


def recover_shape_segm(mask, width, height, pad_info):
    left, right, top, down = pad_info[:4]
    m = np.asarray(mask).astype(f32)
    m = m[:, top:m.shape[1] - down, left:m.shape[2] - right]
    return bilinear_resize(m, height, width)


def plot_all_mask(mask, image, colors, alpha):
    """mask fp32 [K,h,w] (drawing order), image fp32 [h,w,3], colors fp32 [K,3] -> blended image (utils/visualizer.py:95-100)."""
    a = f32(alpha)
    color_mask = ((mask[..., None] * colors[:, None, None, :]).astype(f32) * a).astype(f32)
    alpha_cum = np.cumprod((f32(1) - a * mask).astype(f32), axis=0, dtype=f32)[..., None]
    out = (image * alpha_cum[-1]).astype(f32) + color_mask[0]
    if mask.shape[0] > 1:
        out = out + (color_mask[1:] * alpha_cum[:-1]).astype(f32).sum(axis=0, dtype=f32)
    return out.astype(f32)


def blend_oracle(image, masks, colors, pad_info, alpha):
    h, w = image.shape[:2]
    soft = recover_shape_segm(masks, w, h, pad_info)
    areas = soft.reshape(soft.shape[0], -1).sum(axis=1, dtype=np.float64)
    order = np.argsort(areas, kind='stable')
    return plot_all_mask(soft[order], np.asarray(image, f32), np.asarray(colors, f32)[order], alpha), order, areas
