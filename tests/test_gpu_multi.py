"""The N > 1 path on real GPUs over NCCL (skipped on a single-GPU box; the CPU suite covers the same logic on gloo, tests/test_host.py):
`gather_detections` with the records written by the NMS kernel as the collective's source, issued on the side stream behind the
post-process' `nms_done` event, even and ragged batch splits -- every rank must end up with exactly what one rank computes alone."""
import os
import subprocess
import sys

import pytest
import torch

from tests.common import ROOT

# `multigpu`, not `gpu`: the round-end `-m gpu` run happens on a one-GPU box, where this file is deselected instead of reported as a skip;
# run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m multigpu`
pytestmark = [pytest.mark.multigpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')]

WORKER = r'''
import functools, os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import orienmask_b200 as ob
from orienmask_b200.sharding import gather_detections, shard_bounds
from tests.common import synthetic_heads, post_config
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=dev, **post_config(64, 96, 0.005))
for total in (4, 5):                                   # even split, ragged split (3 + 2)
    heads = synthetic_heads(total, 64, 96, seed=17)
    full = post.apply_padded([(b.to(dev), o.to(dev)) for b, o in heads])          # what one rank computes alone
    lo, hi = shard_bounds(total, rank, world)
    mine = post.apply_padded([(b[lo:hi].to(dev), o[lo:hi].to(dev)) for b, o in heads])
    det, cls, cnt = gather_detections(mine.det, mine.cls, mine.count, packed=mine.packed, total=total, ready=mine.nms_done)
    torch.cuda.synchronize()
    assert torch.equal(cnt.cpu(), full.count.cpu()), (total, cnt, full.count)
    assert torch.equal(det.cpu(), full.det.cpu()) and torch.equal(cls.cpu(), full.cls.cpu()), total
dist.destroy_process_group()
print('rank %%d ok' %% rank)
'''


def test_gather_detections_over_nccl_even_and_ragged(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', str(29600 + os.getpid() % 300), str(script)], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, PYTHONPATH=ROOT))
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'rank 0 ok' in out.stdout and 'rank 1 ok' in out.stdout
