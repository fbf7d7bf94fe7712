"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's detections -> COCO format step.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this.

Follows ``/root/reference/eval/coco_eval.py``: ``_recover_shape_bbox`` :146-188, ``_recover_shape_segm``
:191-205 (crop ``collate_pad`` then ``pad``, flips, ``F.interpolate(bilinear, align_corners=False)``,
``round().to(uint8)``), ``_to_segm_coco_format`` :108-127 and ``_to_bbox_coco_format`` :129-144.

The RLE codec is a third-party dependency that is NOT under /root/reference: ``pycocotools`` (requirements.txt,
unpinned; cocoapi PythonAPI 2.0.x, ``common/maskApi.c``), and it is not installed in this image either.  Its
published algorithm is restated here:
  rleEncode   : walk the mask in column-major (Fortran) order, emit the run lengths of 0s and 1s alternately,
                starting with the (possibly empty) run of 0s; the last run is always emitted.
  rleToString : for count i, x = cnts[i] - (i > 2 ? cnts[i-2] : 0); then 5 bits at a time, low group first:
                c = x & 0x1f; x >>= 5 (arithmetic); more = (c & 0x10) ? x != -1 : x != 0; if more: c |= 0x20;
                emit chr(c + 48).
  rleFrString / rleDecode are the inverses (used for the round-trip checks).
Parity status: pinned by hand-derived known answers and encode->string->decode round trips
(``tests/test_coco_format.py``); the resize part is pinned against ``torch.nn.functional.interpolate``
(mask IoU >= 0.999, exact where no blend lands within one ulp of 0.5) -- see oracle/prep_oracle.py for why
the interpolation has no single bit-exact definition across shapes.
"""
import numpy as np

from .prep_oracle import bilinear_resize

f32 = np.float32


def recover_shape_bbox(bbox, info):
    """eval/coco_eval.py:146-188 on [n,4] fp32 (cx, cy, w, h normalised) -> xywh in original pixels (fp32 ops in order)."""
    b = np.asarray(bbox, dtype=f32)
    bx, by, bw, bh = [b[:, i:i + 1].copy() for i in range(4)]

    def unpad(left, top, right, down, h, w):
        nonlocal bx, by, bw, bh
        nh, nw = h - top - down, w - left - right
        bx = ((bx * f32(w) - f32(left)) / f32(nw)).astype(f32)
        by = ((by * f32(h) - f32(top)) / f32(nh)).astype(f32)
        bw = (bw * f32(w) / f32(nw)).astype(f32)
        bh = (bh * f32(h) / f32(nh)).astype(f32)

    if info.get('collate_pad') is not None:
        left, right, top, down, h, w = info['collate_pad']
        unpad(left, top, right, down, h, w)
    if info.get('pad') is not None:
        top, down, left, right, h, w = info['pad']
        unpad(left, top, right, down, h, w)
    if info.get('hflip', False):
        bx = (f32(1) - bx).astype(f32)
    if info.get('vflip', False):
        by = (f32(1) - by).astype(f32)
    oh, ow = info['height'], info['width']
    bx = ((bx - bw / f32(2)) * f32(ow)).astype(f32)
    by = ((by - bh / f32(2)) * f32(oh)).astype(f32)
    bw = (bw * f32(ow)).astype(f32)
    bh = (bh * f32(oh)).astype(f32)
    return np.concatenate([bx, by, bw, bh], axis=-1)


def crop_window(info, H, W):
    """(top, left, crop_h, crop_w) left by the two crops of eval/coco_eval.py:192-197."""
    top = left = 0
    h, w = H, W
    if info.get('collate_pad') is not None:
        l, r, t, d = info['collate_pad'][:4]
        top += t; left += l; h -= t + d; w -= l + r
    if info.get('pad') is not None:
        t, d, l, r = info['pad'][:4]
        top += t; left += l; h -= t + d; w -= l + r
    return top, left, h, w


def recover_shape_segm(mask, info):
    """eval/coco_eval.py:191-205 on bool/uint8 [K,H,W] -> uint8 [K, height, width]."""
    m = np.asarray(mask).astype(np.uint8)
    top, left, h, w = crop_window(info, m.shape[1], m.shape[2])
    m = m[:, top:top + h, left:left + w]
    if info.get('hflip', False):
        m = m[:, :, ::-1]
    if info.get('vflip', False):
        m = m[:, ::-1, :]
    v = bilinear_resize(m.astype(f32), info['height'], info['width'])
    return (v > f32(0.5)).astype(np.uint8)            # torch.round(): ties to even, v in [0, 1]


def rle_encode(mask):
    """maskApi.c rleEncode of one [h, w] 0/1 array -> run lengths (uint32), column-major, zeros first."""
    flat = np.asarray(mask).astype(np.uint8).reshape(-1, order='F')
    n = flat.size
    if n == 0:
        return np.zeros(1, dtype=np.uint32)
    change = np.flatnonzero(np.concatenate([[flat[0] != 0], flat[1:] != flat[:-1]]))
    pos = np.concatenate([change, [n]]).astype(np.int64)
    return np.diff(np.concatenate([[0], pos])).astype(np.uint32)


def rle_to_string(counts):
    """maskApi.c rleToString."""
    out = bytearray()
    c = [int(v) for v in counts]
    for i, x in enumerate(c):
        if i > 2:
            x -= c[i - 2]
        more = True
        while more:
            ch = x & 0x1f
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(ch + 48)
    return bytes(out)


def rle_from_string(s):
    """maskApi.c rleFrString."""
    if isinstance(s, str):
        s = s.encode('ascii')
    counts = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            ch = s[p] - 48
            x |= (ch & 0x1f) << (5 * k)
            more = bool(ch & 0x20)
            p += 1
            k += 1
            if not more and (ch & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return np.asarray(counts, dtype=np.uint32)


def rle_decode(counts, h, w):
    """maskApi.c rleDecode -> uint8 [h, w]."""
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in counts:
        c = int(c)
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    assert pos == h * w, (pos, h, w)
    return flat.reshape((h, w), order='F')


def to_segm_coco_format(batch_info, detections, cat2label):
    """eval/coco_eval.py:108-127; detections = list of {'bbox' [K,5], 'mask' [K,H,W], 'cls' [K]} numpy arrays."""
    results = []
    for info, det in zip(batch_info, detections):
        if det['bbox'].size == 0:
            continue
        masks = recover_shape_segm(det['mask'], info)
        scores = det['bbox'][:, -1].tolist()
        cats = [int(cat2label[int(c)]) for c in np.asarray(det['cls']).reshape(-1)]
        for m, score, cat in zip(masks, scores, cats):
            rle = {'size': [int(m.shape[0]), int(m.shape[1])], 'counts': rle_to_string(rle_encode(m)).decode('utf-8')}
            results.append({'image_id': info['id'], 'category_id': cat, 'segmentation': rle, 'score': score})
    return results


def to_bbox_coco_format(batch_info, detections, cat2label):
    """eval/coco_eval.py:129-144."""
    results = []
    for info, det in zip(batch_info, detections):
        if det['bbox'].size == 0:
            continue
        xywh = recover_shape_bbox(det['bbox'][:, :4], info).tolist()
        scores = det['bbox'][:, -1].tolist()
        cats = [int(cat2label[int(c)]) for c in np.asarray(det['cls']).reshape(-1)]
        for box, score, cat in zip(xywh, scores, cats):
            results.append({'image_id': info['id'], 'category_id': cat, 'bbox': box, 'score': score})
    return results
