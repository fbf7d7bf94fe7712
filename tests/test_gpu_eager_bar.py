"""The GPU 'bar' of SURVEY §8(d): the reference's forward as stock PyTorch eager ops (cuDNN / cuBLAS) on the same B200,
timed next to the tcgen05 engine on the same weights and batch.  The reference tree is not on the GPU box, so the forward
is oracle/forward_oracle.py -- the same torch calls the reference's nn.Modules make (conv2d, batch_norm, leaky_relu,
nearest interpolate, cat) -- run on CUDA in fp32 (TF32 convs, infer.py's default), fp16, and fp16 channels_last with
cudnn.benchmark=True (infer.py:73-74).  Writes gpurun_out/eager_bar.json; asserts only sanity (heads agree), never speed.
"""
import json
import os

import pytest
import torch

from tests.common import ROOT

pytestmark = pytest.mark.gpu


def _time(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_eager_pytorch_forward_bar():
    import orienmask_b200 as ob
    from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images
    from oracle.forward_oracle import forward_oracle
    B, S = 32, 544
    sd = synthetic_state_dict(0)
    x = synthetic_images(B, S, S, seed=1).cuda()
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    res = {'batch': B, 'size': S, 'gpu': torch.cuda.get_device_name(0)}
    res['engine_fp16_ms'] = _time(lambda: model(x), iters=10)
    ours = [(b.clone(), o.clone()) for b, o in model(x)]

    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        sd32 = {k: v.cuda() for k, v in sd.items()}
        res['eager_fp32_tf32_ms'] = _time(lambda: forward_oracle(sd32, x))
        ref = forward_oracle(sd32, x)

        def half_forward(channels_last):
            # forward_oracle casts to float; re-run its walk in half by monkey-free means: cast inputs, keep BN in half too
            import oracle.forward_oracle as fo
            sdh = {k: (v.half().contiguous(memory_format=torch.channels_last) if (channels_last and v.dim() == 4) else v.half())
                   for k, v in sd32.items() if v.is_floating_point()}
            xh = x.half().contiguous(memory_format=torch.channels_last) if channels_last else x.half()
            orig = torch.Tensor.float
            torch.Tensor.float = lambda t: t            # keep the walk in fp16 (the reference's model.half())
            try:
                return _time(lambda: fo.forward_oracle(sdh, xh))
            finally:
                torch.Tensor.float = orig

        res['eager_fp16_ms'] = half_forward(False)
        res['eager_fp16_channels_last_ms'] = half_forward(True)
    finally:
        torch.backends.cudnn.benchmark = old
    for k in list(res):
        if k.endswith('_ms'):
            res[k.replace('_ms', '_img_s')] = 1e3 * B / res[k]
    res['engine_vs_best_eager'] = min(res['eager_fp32_tf32_ms'], res['eager_fp16_ms'], res['eager_fp16_channels_last_ms']) / res['engine_fp16_ms']
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'eager_bar.json'), 'w'), indent=1)
    print(json.dumps(res))
    # sanity only: both forwards compute the same network
    for (b, o), (rb, ro) in zip(ours, ref):
        assert float((b - rb).norm() / rb.norm()) < 0.03 and float((o - ro).norm() / ro.norm()) < 0.03
