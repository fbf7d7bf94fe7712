"""Batch sharding over ranks and the single collective of the inference path.

Images are independent through forward, decode, NMS and mask assembly (BatchNorm is in eval mode,
the reference loops over images, eval/orienmask_yolo_postprocess.py:75), so a batch is split
contiguously over the ranks with no data-path collective.  The only exchange is an all-gather of
the fixed-size detection records ``[B_local, nms_post, 6]`` fp32 (cx, cy, w, h, score, cls) plus
``[B_local]`` counts -- 76.9 KB per rank at batch 32; masks stay on the GPU that produced them.
The reference itself has no multi-GPU inference (infer.py:69 asserts n_gpu == 1).
"""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous split of ``total`` images: rank r gets [lo, hi); earlier ranks take the remainder."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(images, rank=None, world=None):
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(images.shape[0], rank, world)
    return images[lo:hi]


def pack_records(det, cls, count):
    """[B,K,5] fp32 + [B,K] int64 + [B] int32 -> [B, K*6+1] fp32 rows (count bit-stored as a float value)."""
    B, K, _ = det.shape
    rec = torch.cat([det, cls.to(torch.float32).unsqueeze(-1)], dim=-1).reshape(B, K * 6)
    return torch.cat([rec, count.to(torch.float32).view(B, 1)], dim=1).contiguous()


def unpack_records(packed, K):
    B = packed.shape[0]
    rec = packed[:, :K * 6].reshape(B, K, 6)
    return rec[..., :5].contiguous(), rec[..., 5].to(torch.int64), packed[:, K * 6].to(torch.int32)


_side_streams = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def gather_detections(det, cls, count, group=None, packed=None, total=None, ready=None):
    """All-gather the padded detection records of every rank.

    ``ready``: a CUDA event recorded when the records are final (``PaddedDetections.nms_done``: after the NMS kernel, before the
    mask kernel).  When given, the collective is issued on a side stream that waits for that event only, so it runs under the mask
    kernel instead of behind it on the compute stream; the compute stream joins the side stream before the result is unpacked.

    ``packed``: the [B_local, K*6+1] record rows the NMS kernel wrote (``PaddedDetections.packed``); when given, that buffer
    is the collective's source and nothing is re-packed.
    ``total``: global batch size when it does not divide evenly over the ranks (``shard_bounds`` gives the earlier ranks one
    image more): every rank pads its rows to ``ceil(total / world)`` so that the collective stays fixed-size, and the pad
    rows are dropped on arrival -- the shard sizes follow from ``total`` alone, so no size exchange is needed.
    Returns (det [B_total,K,5], cls [B_total,K], count [B_total]) in rank order, on every rank.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return det, cls, count
    K = det.shape[1]
    if packed is None:
        packed = pack_records(det, cls, count)
    world = dist.get_world_size(group)
    rows = packed.shape[0]
    if total is not None:
        sizes = [hi - lo for lo, hi in (shard_bounds(total, r, world) for r in range(world))]
        if sizes[dist.get_rank(group)] != rows:
            raise ValueError('rank %d holds %d images but shard_bounds(%d, ...) assigns it %d'
                             % (dist.get_rank(group), rows, total, sizes[dist.get_rank(group)]))
        rows = max(sizes)
        if packed.shape[0] < rows:
            packed = torch.cat([packed, packed.new_zeros(rows - packed.shape[0], packed.shape[1])], dim=0)
    packed = packed.contiguous()
    if ready is not None and packed.is_cuda:
        cur = torch.cuda.current_stream(packed.device)
        side = _side_stream(packed.device)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            out = torch.empty(world * rows, packed.shape[1], dtype=packed.dtype, device=packed.device)
            dist.all_gather_into_tensor(out, packed, group=group)
            done = torch.cuda.Event()
            done.record(side)
        packed.record_stream(side)
        out.record_stream(cur)
        cur.wait_event(done)
    else:
        out = torch.empty(world * rows, packed.shape[1], dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out, packed, group=group)
    if total is not None and min(sizes) != rows:
        out = torch.cat([out[r * rows:r * rows + n] for r, n in enumerate(sizes)], dim=0)
    return unpack_records(out, K)
