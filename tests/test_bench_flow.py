"""bench.py's GPU arm walked end to end on CPU stand-ins (no kernel runs): the control flow, the event / stream plumbing of
the device-resident and end-to-end loops, the stage measurements and the JSON contract of the printed line.  The numbers
are meaningless here; the real run happens on the B200 box."""
import contextlib
import io
import json
import os
import sys
import types

import torch


class _Event:
    clock = 0.0

    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        _Event.clock += 1.0
        self.t = _Event.clock

    def elapsed_time(self, other):
        assert self.t is not None and other.t is not None, 'elapsed_time on an event that was never recorded'
        assert other.t > self.t
        return other.t - self.t


class _Stream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def wait_event(self, ev):
        assert ev.t is not None, 'waiting on an event that was never recorded'

    def wait_stream(self, other):
        pass


class _Graph:
    def replay(self):
        pass


B, K = 2, 100


class _Patcher:
    """monkeypatch.setattr for a spawned worker (no pytest fixtures there)."""

    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def install_stand_ins(monkeypatch):
    import bench
    import orienmask_b200 as ob
    from orienmask_b200 import coco_format
    cur = _Stream()
    for name, value in dict(is_available=lambda: True, set_device=lambda d: None, synchronize=lambda *a: None, Event=_Event,
                            Stream=_Stream, current_stream=lambda *a: cur, stream=lambda s: contextlib.nullcontext(),
                            CUDAGraph=_Graph, graph=lambda g: contextlib.nullcontext()).items():
        monkeypatch.setattr(torch.cuda, name, value)
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    monkeypatch.setattr(bench, 'bench_device', lambda t, local: t.device('cpu'))

    class Sampler:
        rows = [1]

        def __init__(self, index):
            pass

        def sparse(self, seconds):
            pass

        def stop(self, t0, t1):
            return {'sm_mhz': 1800.0, 'sm_max_mhz': 1965.0, 'reasons': [], 'samples': 2, 'source': 'stand-in'}
    monkeypatch.setattr(bench, 'ClockSampler', Sampler)

    class Model:
        precision = 'fp16'
        _engines = {0: types.SimpleNamespace(layers=[{'bytes': 10}, {'bytes': 5}])}

        def __init__(self, num_anchors, num_classes):
            pass

        def load_state_dict(self, sd, strict=True):
            assert strict

        def to(self, dev):
            return self

        def eval(self):
            return self

        def __call__(self, x):
            assert x.shape == (B, 3, bench.H, bench.W) and x.dtype == torch.float32
            o = torch.zeros(B, 18, bench.H // 4, bench.W // 4)
            return tuple((torch.zeros(B, 255, bench.H // s, bench.W // s), o[:, 6 * i:6 * i + 6]) for i, s in enumerate((32, 16, 8)))

    class Post:
        nms_pre, nms_post = 400, K

        def __init__(self, **kw):
            assert kw['nms_func'].keywords == {'threshold': 0.5}

        def apply_padded(self, heads):
            out = types.SimpleNamespace(det=torch.zeros(B, K, 5), cls=torch.zeros(B, K, dtype=torch.int64),
                                        count=torch.full((B,), 3, dtype=torch.int32), packed=torch.zeros(B, K * 6 + 1), nms_done=None)
            out.packed[:, -1] = 3.0
            out.to_list = lambda: [{'bbox': torch.zeros(3, 5), 'mask': torch.zeros(3, bench.H, bench.W, dtype=torch.bool),
                                    'cls': torch.zeros(3, dtype=torch.int64)} for _ in range(B)]
            return out

    class Transform:
        def __init__(self, pipeline):
            pass

        def __call__(self, x, out=None):
            assert x.dtype == torch.uint8 and x.shape == (B, bench.H, bench.W, 3)
            return torch.zeros(B, 3, bench.H, bench.W)

    monkeypatch.setattr(ob, 'OrienMaskYOLOFPNPlus', Model)
    monkeypatch.setattr(ob, 'OrienMaskYOLOPostProcess', Post)
    monkeypatch.setattr(ob, 'FastCOCOTransform', Transform)
    monkeypatch.setattr(coco_format, 'encode_masks', lambda masks, counts, infos: [[{'size': [1, 1], 'counts': 'ab'}] * n for n in counts])
    from orienmask_b200 import synthetic
    monkeypatch.setattr(synthetic, 'synthetic_state_dict', lambda seed=0: {})
    return bench


def test_gpu_arm_flow_and_json_contract(monkeypatch, capsys):
    bench = install_stand_ins(monkeypatch)
    monkeypatch.setattr(bench, 'make_cpu_reference', lambda: ({}, None))                  # the CPU leg itself: test_host.py (reference arm)
    monkeypatch.setattr(bench, 'cpu_reference_step', lambda n, sd, post, threads: 0.5 * n)
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--steps', '3', '--warmup', '1', '--batch', str(B)])
    for key in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        monkeypatch.delenv(key, raising=False)
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'stages', 'step_ms'):
        assert key in line, key
    assert line['steps'] == 3 and line['warmup'] == 3 and line['n_gpus'] == 1 and line['scaling'] == 'weak'       # W >= 3 is enforced
    assert line['metric'] == bench.METRIC and line['unit'] == 'images/sec' and line['vs_baseline'] is None
    assert 'workload' in line['config'] and 'model' not in line['config']
    e2e = line['e2e']
    assert e2e['value'] > 0 and e2e['wall_value'] > 0 and e2e['h2d_bytes_per_step'] == B * bench.H * bench.W * 3
    assert e2e['d2h_bytes_per_step'] == B * K * 5 * 4 + B * 4
    rf = line['roofline']
    assert rf['bound'] == 'tensor' and rf['unit'] == 'TFLOP/s' and abs(rf['frac'] - rf['achieved'] / rf['peak']) < 1e-12
    assert rf['traffic_algorithmic'] == 15
    assert set(line['stages']) == {'preprocess', 'postprocess', 'coco_format'}
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['unit'] == 'images/sec' and abs(cb['value'] - 2.0) < 1e-9 and cb['cores'] >= 1 and 'sample' in cb
    assert cb['os_cpu_count'] == cb['cores'] and cb['torch_threads'] >= 1 and 'cpu_model' in cb
    pp = line['stages']['postprocess']
    assert pp['bytes'] == B * (255 * (17 * 17 + 34 * 34 + 68 * 68) + 18 * 136 * 136) * 4 + 2 * 3 * bench.H * bench.W


def _rank_worker(rank, world, port, q):
    """One rank of `torchrun bench.py --gpus 2` on stand-ins: gloo instead of NCCL, CPU tensors instead of device tensors."""
    import io
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    bench = install_stand_ins(_Patcher())
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, device_id=None: real_init('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    sys.argv = ['bench.py', '--gpus', str(world), '--steps', '3', '--warmup', '3', '--batch', str(B)]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    q.put((rank, buf.getvalue()))


def test_two_rank_flow_on_gloo():
    """The N > 1 path of bench.py (barriers, max-over-ranks reductions, the detection all-gather, rank 0 alone printing)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert results[1].strip() == ''                                   # only rank 0 prints
    lines = [l for l in results[0].splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['n_gpus'] == 2 and line['config']['global_batch'] == 2 * B and line['config']['parallelism'] == 'dp2'
    assert line['scaling'] == 'weak' and 'cpu_baseline' not in line    # the CPU baseline is an N = 1 leg
    assert line['e2e']['h2d_bytes_per_step'] == 2 * B * 544 * 544 * 3 and line['e2e']['d2h_bytes_per_step'] == 2 * B * K * 5 * 4 + 2 * B * 4
    assert line['e2e']['value'] > 0 and line['value'] > 0
