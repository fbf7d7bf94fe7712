#!/usr/bin/env python
"""Headline benchmark: images/sec of the OrienMask hot path at 544x544, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = DarkNet-53+FPNPlus forward -> decode/select -> class-wise NMS -> mask assembly on one
synthetic batch (random-init weights of the reference architecture, uniform-noise images; no
datasets or checkpoints are reachable offline).  N > 1: one process per GPU under torchrun, the
batch dimension is sharded (weak scaling: 32 images per rank) and the only collective is the NCCL
all-gather of the padded detection records.  Prints ONE JSON line on rank 0 (contract in the task
statement): `value` is device-resident throughput, `e2e` goes through the public API from pinned
host memory, `roofline` is the conv engine against the measured tensor peak, `cpu_baseline` /
`--impl reference` time the CPU restatement of the reference (oracle/) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 544
BATCH = 32
GFLOP_PER_IMAGE = 173.845           # 2 * 86.9226 GMAC over the 90 convolutions
METRIC = 'images/sec @544x544 bs32 (forward + decode + NMS + masks)'
ANCHORS = [[12, 16], [19, 36], [40, 28], [36, 75], [76, 55], [72, 146], [142, 110], [192, 243], [459, 401]]
ANCHOR_MASK = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]


def post_kwargs():
    return dict(grid_size=[[H // s, W // s] for s in (32, 16, 8)], image_size=[H, W], anchors=ANCHORS,
                anchor_mask=ANCHOR_MASK, num_classes=80, conf_thresh=0.005, nms_pre=400, nms_post=100, orien_thresh=0.3)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return dict(tflops=float(d.get('bf16_tflops_sustained') or d.get('bf16_tflops')), hbm=float(d.get('hbm_gbs')),
                        source='MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)')
        except Exception:
            pass
    return dict(tflops=1400.0, hbm=6650.0, source='fallback of B200_PROFILING.md (1.4 PFLOP/s sustained, 6.65 TB/s)')


def measured_traffic():
    """DRAM bytes (read + write) of the conv-engine launches of one bs-32 step from the committed ncu pass
    (tools/traffic_report.py -> profiles/r01_traffic.json); None when no capture is committed."""
    for name in ('r02_traffic.json', 'r01_traffic.json'):
        try:
            f = json.load(open(os.path.join(ROOT, 'profiles', name)))['conv_engine']
            return {'bytes': f['dram_read'] + f['dram_write'], 'launches': f['launches'], 'file': 'profiles/' + name}
        except Exception:
            continue
    return None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.

    In-process NVML (nvidia_ml_py) from a thread.  Every driver query stalls kernel submission (an `nvidia-smi -lms` child
    for tens of milliseconds per poll, an NVML field query for a few): with 20 ms polling a 10-step timed region (65 ms)
    read anything between 6.7 and 20 ms per step while the un-sampled end-to-end loop stayed within 2 %.  So the sampler
    polls every 20 ms during the warm-up (same workload, immediately before) and, once `sparse(seconds)` is called with the
    expected length of the timed region, takes only two samples inside it, in its second half.  nvidia-smi is the fallback."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag, self.period, self.next_at = [], None, None, False, 0.02, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            if self.next_at is not None:                      # sparse mode: wait for the scheduled instant
                if time.time() < self.next_at:
                    time.sleep(0.004)
                    continue
                self.next_at += self.period
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                bits = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append((time.time(), mhz, bits))
            except Exception:
                pass
            if self.next_at is None:
                time.sleep(self.period)

    def sparse(self, region_seconds):
        """Two samples, at ~55 % and ~85 % of the timed region: by then the host is several steps ahead of the GPU, so a driver
        query that blocks kernel submission for a few milliseconds no longer idles the device."""
        self.period = max(0.02, 0.3 * region_seconds)
        self.next_at = time.time() + 0.03 + 0.55 * region_seconds

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self.stop_flag = True
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
            if not rows:
                return None
            bits = 0
            for r in rows:
                bits |= r[2]
            return {'sm_mhz': statistics.median(r[1] for r in rows), 'sm_max_mhz': self.max_mhz,
                    'reasons': [name for name, b in self.BITS if bits & b], 'samples': len(rows), 'source': 'nvml'}
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
        if not rows:
            return None
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': float(rows[0][1]) if rows[0][1].isdigit() else None,
                'reasons': reasons, 'samples': len(rows), 'source': 'nvidia-smi'}


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(n_images, sd, post, threads):
    """One pass of the CPU restatement of the reference (oracle/) over n_images; returns seconds."""
    import torch
    from oracle.forward_oracle import forward_oracle
    from orienmask_b200.synthetic import synthetic_images
    torch.set_num_threads(threads)
    x = synthetic_images(n_images, H, W, seed=1)
    t0 = time.perf_counter()
    heads = forward_oracle(sd, x)
    t1 = time.perf_counter()
    post([(b.numpy(), o.numpy()) for b, o in heads])
    t2 = time.perf_counter()
    CPU_SPLIT[0] += t1 - t0                 # forward / post-process seconds of the CPU arm, reported next to its throughput
    CPU_SPLIT[1] += t2 - t1
    return t2 - t0


CPU_SPLIT = [0.0, 0.0]


def cpu_split(images):
    """Per-image forward / post-process milliseconds accumulated by cpu_reference_step since the last call."""
    fwd, post = CPU_SPLIT
    CPU_SPLIT[0] = CPU_SPLIT[1] = 0.0
    return {'forward_ms_per_image': 1e3 * fwd / images, 'postprocess_ms_per_image': 1e3 * post / images}


def cpu_description():
    """Host the CPU arm ran on: model string, os.cpu_count(), torch intra-op threads."""
    import torch
    model = None
    try:
        for line in open('/proc/cpuinfo'):
            if line.lower().startswith('model name'):
                model = line.split(':', 1)[1].strip()
                break
    except OSError:
        pass
    return {'cpu_model': model, 'os_cpu_count': os.cpu_count(), 'torch_threads': torch.get_num_threads()}


def make_cpu_reference():
    from oracle.post_oracle import PostProcessOracle
    from orienmask_b200.synthetic import synthetic_state_dict
    kw = post_kwargs()
    post = PostProcessOracle(kw['grid_size'], kw['image_size'], kw['anchors'], kw['anchor_mask'], kw['num_classes'],
                             conf_thresh=kw['conf_thresh'], nms_threshold=0.5, nms_pre=kw['nms_pre'],
                             nms_post=kw['nms_post'], orien_thresh=kw['orien_thresh'])
    return synthetic_state_dict(0), post


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sd, post = make_cpu_reference()
    sample = 2
    for _ in range(args.warmup):
        cpu_reference_step(sample, sd, post, threads)
    cpu_split(1)
    t = sum(cpu_reference_step(sample, sd, post, threads) for _ in range(args.steps))
    v = sample * args.steps / t
    split = cpu_split(sample * args.steps)
    desc = {'value': v, 'unit': 'images/sec', 'cores': threads, 'kind': 'port',
            'sample': '%d images of 544x544 per step (forward via torch CPU/oneDNN fp32 + numpy/C post-process)' % sample}
    desc.update(cpu_description())
    desc.update(split)
    print(json.dumps({'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/sec', 'n_gpus': args.gpus,
                      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps,
                      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': {'workload': 'bs=%d %dx%d per GPU: DarkNet-53+FPNPlus forward + decode + batched NMS + mask assembly' % (BATCH, H, W),
                                 'sample': 'each step is a bounded CPU sample of %d images of that workload' % sample,
                                 'weights': 'synthetic_state_dict(seed 0)', 'precision': 'fp32 (torch CPU / oneDNN)'},
                      'cpu_baseline': desc,
                      'e2e': {'value': v, 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def measure_stages(torch, ob, transform, host_u8, out, dev, iters=10, post=None, heads=None):
    """The two neighbours of the path, each timed alone with CUDA events against the HBM roofline:
    pre-process (uint8 HWC -> fp32 NCHW) and detections -> COCO RLE (masks -> run-length strings for 480x640 originals)."""
    from orienmask_b200.coco_format import encode_masks
    from orienmask_b200 import _lib
    hbm = measured_peaks()['hbm']
    x = host_u8.to(dev)
    res = {}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    # replayed from a CUDA graph of 10 launches: the kernel takes tens of microseconds, less than a Python call
    buf = transform(x)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        transform(x, out=buf)
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(10):
            transform(x, out=buf)
    us = timed(graph.replay) / 10
    nbytes = x.numel() + x.shape[0] * 3 * H * W * 4
    res['preprocess'] = {'us': us, 'bytes': nbytes, 'GBps': nbytes / us / 1e3, 'frac_of_hbm': nbytes / us / 1e3 / hbm,
                         'what': 'uint8 HWC [%d,%d,%d,3] -> fp32 NCHW, one prep_kernel launch' % tuple(x.shape[:3])}
    dets = out.to_list()
    counts = [int(d['bbox'].shape[0]) for d in dets]
    if post is not None and heads is not None:
        # the post-process half of the headline step alone (decode/select + NMS + masks: 5 launches), against the HBM roofline.
        # Algorithmic bytes: every head value read once, one mask byte written per (instance, pixel).
        us = timed(lambda: post.apply_padded(heads))
        nbytes = sum(int(b.numel()) * 4 for b, _ in heads) + int(heads[0][1].numel()) * 3 * 4 + sum(counts) * H * W
        res['postprocess'] = {'us': us, 'bytes': nbytes, 'GBps': nbytes / us / 1e3, 'frac_of_hbm': nbytes / us / 1e3 / hbm,
                              'what': 'decode + confidence filter + top-%d + class-wise NMS + top-%d + mask assembly for %d images '
                                      '(%d instances); the mask kernel is write-only (measured write ceiling 3.9 TB/s, DESIGN finding 12)'
                                      % (post.nms_pre, post.nms_post, len(counts), sum(counts))}
    infos = [{'id': i, 'height': 480 * H // 544, 'width': 640 * W // 544, 'collate_pad': [0, 0, 0, 0, H, W]} for i in range(len(dets))]
    masks = [d['mask'] for d in dets]
    lib = _lib.lib()
    lib.om_launch_count_reset()
    t0 = time.perf_counter()
    enc = encode_masks(masks, counts, infos)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = int(lib.om_launch_count())
    # kernel alone: CUDA events around repeated calls of the same launch (through encode_masks' C-ABI call)
    ev = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        encode_masks(masks, counts, infos)
        e1.record()
        torch.cuda.synchronize()
        ev.append(e0.elapsed_time(e1))
    nbytes = sum(counts) * H * W
    text = sum(len(e['counts']) for im in enc for e in im)
    res['coco_format'] = {'ms_host_to_strings': min(ev), 'first_call_ms': wall_ms, 'instances': sum(counts), 'mask_bytes_read': nbytes,
                          'rle_text_bytes': text, 'GBps_end_to_end': nbytes / min(ev) / 1e6, 'launches': launches,
                          'what': 'bool masks [K,%d,%d] -> crop/resize to the 4:3 original -> round -> column-major RLE -> COCO strings on the '
                                  'device (mask_rle_kernel), strings to host' % (H, W)}
    return res


def _timed_steps(torch, fn, steps, warmup):
    """ms per call of fn(i) over `steps` calls after `warmup`, CUDA events on the launching stream."""
    keep = None
    for i in range(warmup):
        keep = fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        keep = fn(i)
    e1.record()
    torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / steps


def measure_extras(torch, ob, model, post, resident, dev, peaks):
    """Secondary, driver-visible measurements on rank 0 at N = 1, after the headline region (each a few hundred milliseconds of GPU time):

    * `parity_mode`   the same step with the conv engine in its tensor-core parity mode (the configuration whose outputs meet the
                      north-star tolerances end to end: tests/test_gpu_forward.py::test_end_to_end_parity_vs_oracle_544);
    * `reference_api` the headline step through `post(heads)` -- the reference's return type (trimmed per-image dicts: one host
                      synchronisation per step on the per-image counts) instead of the padded device-resident result;
    * `latency_bs1`   BASELINE config 2: batch 1, forward + post-process latency, median of 200 iterations (CUDA graph of the forward);
    * `config5_960`   BASELINE config 5: batch 8 at 960x960."""
    import functools
    import statistics
    from orienmask_b200.arch import macs_per_image
    from orienmask_b200.synthetic import synthetic_images
    res = {}

    def fwd_ms_of(m, x, iters=5):
        return _timed_steps(torch, lambda i: m(x), iters, 2)

    # ---- reference API: post() returns the reference's list of dicts (host sync on the counts inside) ----
    ms = _timed_steps(torch, lambda i: post(model(resident[i % 2])), 10, 3)
    res['reference_api'] = {'value': BATCH * 1e3 / ms, 'unit': 'images/sec', 'ms_per_step': ms,
                            'what': 'model(x) -> postprocess(heads) returning the per-image dicts of trimmed tensors '
                                    '(eval/orienmask_yolo_postprocess.py:124,166); the host waits for the NMS kernel only, the mask '
                                    'kernel overlaps the slicing'}
    # ---- parity mode ----
    model.precision = 'parity'
    ms = _timed_steps(torch, lambda i: post.apply_padded(model(resident[i % 2])), 6, 3)
    f_ms = fwd_ms_of(model, resident[0])
    ach = BATCH * GFLOP_PER_IMAGE / f_ms
    res['parity_mode'] = {'value': BATCH * 1e3 / ms, 'unit': 'images/sec', 'ms_per_step': ms, 'forward_ms': f_ms, 'steps': 6, 'warmup': 3,
                          'dtype': 'f16x2 (hi+lo pairs, three tcgen05 MMAs per product, fp32 accumulate)',
                          'roofline': {'bound': 'tensor', 'achieved': ach, 'issued': 3 * ach, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                                       'frac': ach / peaks['tflops'], 'frac_issued': 3 * ach / peaks['tflops'],
                                       'note': 'achieved = algorithmic FLOPs (173.845 GFLOP/image) / forward time; issued = the 3x tensor work '
                                               'the split-precision products cost'},
                          'parity': 'north-star tolerances end to end (boxes/scores 1e-3, mask IoU 0.999, kept sets identical up to listed '
                                    'margin-limited pairs): tests/test_gpu_forward.py::test_end_to_end_parity_vs_oracle_544'}
    model.precision = 'fp16'
    model._engines = {}
    torch.cuda.empty_cache()
    # ---- config 2: bs 1 latency ----
    x1 = resident[0][:1].contiguous()
    lat = {}
    for mode in ('eager', 'graph'):
        model.use_cuda_graph = mode == 'graph'
        for _ in range(20):
            post.apply_padded(model(x1))
        torch.cuda.synchronize()
        fwd, tot = [], []
        for _ in range(200):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            hd = model(x1)
            e[1].record()
            post.apply_padded(hd)
            e[2].record()
            torch.cuda.synchronize()
            fwd.append(e[0].elapsed_time(e[1]))
            tot.append(e[0].elapsed_time(e[2]))
        lat[mode] = {'forward_ms_median': statistics.median(fwd), 'total_ms_median': statistics.median(tot),
                     'total_ms_p90': sorted(tot)[int(0.9 * len(tot))]}
        model._engines = {}
    model.use_cuda_graph = False
    best = min(lat, key=lambda k: lat[k]['total_ms_median'])
    res['latency_bs1'] = {'value': lat[best]['total_ms_median'], 'unit': 'ms', 'higher_is_better': False, 'mode': best, 'iters': 200, 'warmup': 20,
                          'images_per_s': 1e3 / lat[best]['total_ms_median'], **lat,
                          'what': 'BASELINE config 2: bs=1 544x544, fp16 forward + decode + NMS + mask assembly, one synchronised iteration at a time'}
    torch.cuda.empty_cache()
    # ---- config 5: 960x960, bs 8 ----
    S, B5 = 960, 8
    post5 = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=dev,
                                        grid_size=[[S // s, S // s] for s in (32, 16, 8)], image_size=[S, S], anchors=ANCHORS,
                                        anchor_mask=ANCHOR_MASK, num_classes=80, conf_thresh=0.005, nms_pre=400, nms_post=100, orien_thresh=0.3)
    xs = [synthetic_images(B5, S, S, seed=11 + i).to(dev) for i in range(2)]
    ms = _timed_steps(torch, lambda i: post5.apply_padded(model(xs[i % 2])), 10, 3)
    f_ms = fwd_ms_of(model, xs[0])
    gflop = 2e-9 * macs_per_image(S, S)[0]
    res['config5_960'] = {'value': B5 * 1e3 / ms, 'unit': 'images/sec', 'ms_per_step': ms, 'forward_ms': f_ms, 'steps': 10, 'warmup': 3,
                          'roofline': {'bound': 'tensor', 'achieved': B5 * gflop / f_ms, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                                       'frac': B5 * gflop / f_ms / peaks['tflops'], 'algorithmic': '%.3f GFLOP/image x %d' % (gflop, B5)},
                          'what': 'BASELINE config 5: bs=8 960x960 (stride-4 map 240x240, 56 700 predictions), fp16 forward + post-process'}
    model._engines = {}
    torch.cuda.empty_cache()
    return res


def bench_device(torch, local):
    """The rank's GPU (a seam for tests/test_bench_flow.py, which walks main() on stand-ins without a GPU)."""
    return torch.device('cuda', local)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--precision', default='fp16', choices=['fp16', 'parity', 'fp32'],
                    help="conv engine of the headline line: fp16 (production), parity (tensor cores on fp16 hi+lo pairs, fp32-grade), fp32 (FFMA)")
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary measurements (parity mode, bs 1 latency, 960x960 bs 8, reference API)')
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ncu-range', action='store_true', help='cudaProfilerStart/Stop around the timed device-resident loop '
                    '(ncu --profile-from-start off): the launch list of exactly the timed steps; numbers printed under ncu are not bench values')
    ap.add_argument('--size', type=int, default=544, help='square input size (960 with --batch 8 = config 5); the headline metric is 544')
    args = ap.parse_args()
    global H, W, GFLOP_PER_IMAGE
    if args.size != H:
        from orienmask_b200.arch import macs_per_image
        H = W = args.size
        GFLOP_PER_IMAGE = 2e-9 * macs_per_image(H, W)[0]
        args.no_cpu_baseline = True
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import functools
    import orienmask_b200 as ob
    from orienmask_b200 import _lib
    from orienmask_b200.sharding import gather_detections
    from orienmask_b200.synthetic import synthetic_state_dict, synthetic_images

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the product path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = bench_device(torch, local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG', 'WARN')        # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    model = ob.OrienMaskYOLOFPNPlus(3, 80)
    model.load_state_dict(synthetic_state_dict(0), strict=True)
    model.precision = args.precision
    model = model.to(dev).eval()
    post = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.5), device=dev, **post_kwargs())
    # two distinct resident batches, alternated (per-step activation traffic is >> the 126 MB L2 anyway).  The host side
    # holds what cv2.imread yields (uint8 HWC, infer.py:147); the device-resident arm holds the transformed model input.
    transform = ob.FastCOCOTransform([dict(type='Resize', size=(H, W)), dict(type='Normalize', mean=(0, 0, 0), std=(255, 255, 255))])
    host = [(synthetic_images(B, H, W, seed=1 + rank * 2 + i) * 255).round().clamp(0, 255).to(torch.uint8)
            .permute(0, 2, 3, 1).contiguous().pin_memory() for i in range(2)]
    resident = [transform(h.to(dev)) for h in host]
    lib = _lib.lib()

    def step(x):
        heads = model(x)
        out = post.apply_padded(heads)
        det, cls, cnt = gather_detections(out.det, out.cls, out.count, packed=out.packed, ready=out.nms_done)
        return out, det, cls, cnt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler is started BEFORE the warm-up: NVML initialisation stalls kernel launches for tens of milliseconds
    sampler = ClockSampler(local) if rank == 0 else None
    # Warm-up with the SAME object lifetimes as the timed loop: `heads` / `out` of step i stay alive until step i+1 has
    # allocated its own, so the caching allocator must hold two sets of head / mask tensors (~1.3 GB each).  A warm-up that
    # dropped each result at once left one set cached, and the second step of the timed region paid a synchronising
    # cudaMalloc (one 20-130 ms step in an otherwise 6.3 ms series).
    heads = out = None
    for i in range(args.warmup):
        heads = model(resident[i % 2])
        out = post.apply_padded(heads)
        gather_detections(out.det, out.cls, out.count, packed=out.packed, ready=out.nms_done)
    barrier()
    if sampler is not None:
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 3.0:
            time.sleep(0.05)
    for i in range(2):                       # every rank (the step contains a collective): absorbs rank 0's wait above
        heads = model(resident[i % 2])
        out = post.apply_padded(heads)
        gather_detections(out.det, out.cls, out.count, packed=out.packed, ready=out.nms_done)
    barrier()

    # ---- device-resident timing -------------------------------------------------------------
    fwd_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_end = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    lib.om_launch_count_reset()
    if sampler is not None:
        sampler.sparse(args.steps * 0.007)
        time.sleep(0.03)
    # no garbage collection inside the timed regions: a generation-2 pass of the interpreter (tens of milliseconds with torch's object
    # graph) in the first steps, before the host is ahead of the GPU, showed up as one 15 ms step in an otherwise 6.4 ms series
    import gc
    gc.collect()
    gc.disable()
    barrier()
    if args.ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        fwd_ev[i][0].record()
        heads = model(resident[i % 2])
        fwd_ev[i][1].record()
        out = post.apply_padded(heads)
        gather_detections(out.det, out.cls, out.count, packed=out.packed, ready=out.nms_done)
        step_end[i].record()
    e1.record()
    barrier()
    if args.ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    t1 = time.time()
    launches = int(lib.om_launch_count())
    clocks = sampler.stop(t0, t1) if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    fwd_ms = statistics.mean(a.elapsed_time(b) for a, b in fwd_ev)
    per_step = [(e0 if i == 0 else step_end[i - 1]).elapsed_time(step_end[i]) for i in range(args.steps)]
    k_avg = int(out.count.float().mean().item())

    # ---- end to end through the public API from pinned host memory -----------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(host[0], device=dev) for _ in range(2)]
    rec_host = torch.empty(B * world, 100, 5, dtype=torch.float32).pin_memory()
    cnt_host = torch.empty(B * world, dtype=torch.int32).pin_memory()
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n, span=None, prep_on_copy_stream=True):
        """Double-buffered: while step i runs on the current stream, the side stream copies batch i+1 from pinned host memory and (by
        default) runs its pre-process kernel there too -- FastCOCOTransform is one memory-bound launch that fits into the idle SM time of
        the forward's wave tails; on the current stream it is 60-70 us in front of every forward."""
        cur = torch.cuda.current_stream()
        if span is not None:                          # device-timed: the side stream's first copy starts after the start event
            span[0].record(cur)
            copy_stream.wait_event(span[0])
        xs = [None, None]

        def stage(slot, batch):                       # on the side stream: H2D copy (+ pre-process) of `batch` into slot
            with torch.cuda.stream(copy_stream):
                bufs[slot].copy_(host[batch % 2], non_blocking=True)
                if prep_on_copy_stream:
                    xs[slot] = transform(bufs[slot])  # infer.py:149: permute + resize + normalise (one kernel)
                    if xs[slot].is_cuda:
                        xs[slot].record_stream(cur)
                ready[slot].record(copy_stream)

        stage(0, 0)
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                if i >= 1:
                    copy_stream.wait_event(freed[1 - s])
                stage(1 - s, i + 1)
            cur.wait_event(ready[s])
            x = xs[s] if prep_on_copy_stream else transform(bufs[s])
            _, det, cls, cnt = step(x)
            freed[s].record(cur)                      # bufs[s] / xs[s] may be overwritten once this step's forward has consumed them
            rec_host.copy_(det, non_blocking=True)
            cnt_host.copy_(cnt, non_blocking=True)
        if span is not None:
            span[1].record(cur)                       # after the last device->host copy of the results
        torch.cuda.synchronize()

    def timed_e2e(prep_on_copy_stream):
        e2e_loop(6, None, prep_on_copy_stream)
        gc.collect()
        barrier()
        span = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        w0 = time.perf_counter()
        e2e_loop(args.steps, span, prep_on_copy_stream)
        barrier()
        # [device seconds between the events, host wall seconds incl. the final synchronize]; the device time is the reported one
        t = torch.tensor([span[0].elapsed_time(span[1]) * 1e-3, time.perf_counter() - w0], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in t.tolist())

    e2e_serial_s, _ = timed_e2e(False)               # pre-process on the compute stream, in front of every forward
    e2e_s, e2e_wall_s = timed_e2e(True)              # pre-process of batch i+1 on the copy stream, under step i (the reported number)
    gc.enable()

    stages = measure_stages(torch, ob, transform, host[0], out, dev, post=post, heads=heads) if rank == 0 else None

    if rank == 0:
        peaks = measured_peaks()
        value = world * B * args.steps / (total_ms * 1e-3)
        achieved = B * GFLOP_PER_IMAGE / fwd_ms          # GFLOP / ms == TFLOP/s
        traffic = measured_traffic() if (B == BATCH and H == 544) else None
        eng = next(iter(model._engines.values()))
        algo_bytes = int(sum(l['bytes'] for l in eng.layers))     # per layer: activations in + out (+ addends) + weights, fp16
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/sec', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp16': 'f16', 'parity': 'f16x2 (hi+lo pairs, fp32-grade)', 'fp32': 'f32'}[args.precision], 'data': 'synthetic',
            'config': {'workload': 'bs=%d %dx%d per GPU: DarkNet-53+FPNPlus forward + decode + batched NMS + mask assembly' % (B, H, W),
                       'global_batch': B * world, 'parallelism': 'dp%d' % world, 'weights': 'synthetic_state_dict(seed 0)',
                       'l2': 'two alternating resident batches; per-step activation traffic (>10 GB) >> 126 MB L2',
                       'precision': {'fp16': 'fp16 storage / fp32 accumulate convs, fp32 heads + post-process',
                                     'parity': 'fp16 hi+lo pairs on the tensor cores (3 MMAs per product), fp32 accumulate, fp32 heads + post-process',
                                     'fp32': 'fp32 (FFMA)'}[args.precision],
                       'avg_instances_per_image': k_avg},
            'e2e': {'value': world * B * args.steps / e2e_s, 'unit': 'images/sec',
                    'h2d_bytes_per_step': int(host[0].numel() * host[0].element_size()) * world,
                    'd2h_bytes_per_step': int(rec_host.numel() * 4 + cnt_host.numel() * 4),
                    'wall_value': world * B * args.steps / e2e_wall_s,
                    'value_prep_on_compute_stream': world * B * args.steps / e2e_serial_s,
                    'note': 'pinned uint8 HWC images (cv2 layout) -> FastCOCOTransform -> model() -> postprocess -> detection records + '
                            'counts to host; the copy AND the pre-process kernel of batch i+1 run double-buffered on a side stream under step i '
                            '(value_prep_on_compute_stream: the same loop with the pre-process on the compute stream in front of every forward); '
                            'value = CUDA events from before the first host->device copy to after the last device->host copy, max over ranks; '
                            'wall_value = host clock around the same loop'},
            'step_ms': {'min': min(per_step), 'median': statistics.median(per_step), 'max': max(per_step),
                        'argmax': per_step.index(max(per_step))},
            'gpu_launches': launches,
            'clocks': clocks,
            'stages': stages,
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tflops'], 'unit': 'TFLOP/s',
                         'frac': achieved / peaks['tflops'], 'traffic': traffic['bytes'] if traffic else None,
                         'traffic_algorithmic': algo_bytes,
                         'traffic_note': ('ncu dram__bytes_read+write summed over the %d conv-engine launches of one step '
                                          '(%s); traffic_algorithmic = the un-fused per-layer in + out + weight bytes of the same launches '
                                          '(measured < algorithmic: consecutive layers hit in the 126 MB L2)'
                                          % (traffic['launches'], traffic['file'])) if traffic else None,
                         'kernel': 'conv engine = the model forward: %d launches (conv_tc2_kernel, tcgen05 cta_group::2; fused DarkNet block; '
                                   'TMA-fed tensor-core stem, fused with the stride-2 layer behind it in the fp16 engine), %.3f ms of %.3f ms per step' % (len(eng.layers), fwd_ms, total_ms / args.steps),
                         'algorithmic': '%.3f GFLOP/image x %d images per forward' % (GFLOP_PER_IMAGE, B),
                         'peak_source': peaks['source']},
        }
        if world == 1 and not args.no_extras and args.precision == 'fp16' and H == 544 and B == BATCH:
            del heads, out
            line.update(measure_extras(torch, ob, model, post, resident, dev, peaks))
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sd, cpost = make_cpu_reference()
            cpu_reference_step(1, sd, cpost, threads)
            n_img, reps = 4, 3
            cpu_split(1)
            t = sum(cpu_reference_step(n_img, sd, cpost, threads) for _ in range(reps))
            line['cpu_baseline'] = {'value': n_img * reps / t, 'unit': 'images/sec', 'cores': threads, 'kind': 'port',
                                    'sample': '%d passes over %d images of 544x544 (oracle: torch CPU fp32 forward + numpy/C post-process)' % (reps, n_img)}
            line['cpu_baseline'].update(cpu_description())
            line['cpu_baseline'].update(cpu_split(n_img * reps))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
