"""``OrienMaskYOLOFPNPlus`` with the reference interface, executed by the sm_100a convolution engine.

Interface mirrored (``/root/reference/model/orienmask_yolo_fpnplus.py:8-37,74-90``): same
constructor arguments, the same 524 state-dict keys (so ``load_state_dict(strict=True)`` of a
reference checkpoint works), ``forward(x) -> ((bbox32, orien32), (bbox16, orien16), (bbox8, orien8))``
with fp32 NCHW tensors.  The modules below only *hold* parameters under the reference's names; the
arithmetic is a static schedule of C-ABI kernels over a preallocated buffer plan:

* BN folded into fp16 (or fp32, parity mode) weights once per weight version;
* every ConvBNLeaky = one kernel (bias + LeakyReLU in the epilogue), residual adds fused into the
  3x3 of each DarkNet block, in place;
* ``cat([nearest_up(route), x])`` followed by a 1x1 conv is evaluated as
  ``W_x * x + nearest_up(W_r * route)``: the low-resolution product is a small fp32 "partial" that the
  high-resolution kernel adds in its epilogue, so neither the up-sampled nor the concatenated tensor
  ever exists (both are exact re-associations of the reference's sum);
* inference only (BatchNorm always uses running statistics); CUDA only -- no fallback.
"""
import ctypes
import math
import os

import torch
import torch.nn as nn

from . import _lib
from .arch import conv_specs, STAGE_BLOCKS, STAGE_CHANNELS


class _Params(nn.Module):
    """Parameter holder; never called."""

    def forward(self, *a, **k):
        raise RuntimeError('parameter holder')


def _holder(root, dotted):
    mod = root
    for part in dotted.split('.'):
        if part not in mod._modules:
            mod.add_module(part, _Params())
        mod = mod._modules[part]
    return mod


def _invalidate_after_load(module, incompatible_keys):
    """load_state_dict post-hook (also fires when a PARENT module loads: PyTorch recurses through _load_from_state_dict and never calls
    this module's own load_state_dict override)."""
    module._invalidate()


class OrienMaskYOLOFPNPlus(nn.Module):
    _plus = True          # False in OrienMaskYOLO: no skip convolutions, neck4 reads cat[up2(route8(neck8)), x4]

    def __init__(self, num_anchors, num_classes, pretrained=None, freeze_backbone=False, backbone_batchnorm_eval=False):
        super().__init__()
        self.num_anchors, self.num_classes = num_anchors, num_classes
        # 'fp16' (tcgen05, production) | 'parity' (tcgen05 on fp16 hi + lo pairs, three MMAs per product: fp32-grade results on the
        # tensor cores) | 'fp32' (FFMA on the CUDA cores, the bring-up reference of both)
        self.precision = os.environ.get('ORIENMASK_B200_PRECISION', 'fp16')
        # replay the forward's ~95 launches as one CUDA graph (small-batch latency; outputs are overwritten by the next call)
        self.use_cuda_graph = os.environ.get('ORIENMASK_B200_GRAPH', '0') == '1'
        self._specs = conv_specs(num_anchors, num_classes, self._plus)
        for s in self._specs:
            fan_in = s.cin * s.k * s.k
            w = torch.empty(s.cout, s.cin, s.k, s.k)
            nn.init.kaiming_uniform_(w, a=math.sqrt(5))
            if s.kind == 'cbl':
                conv = _holder(self, s.prefix + '.conv_block.0')
                conv.weight = nn.Parameter(w)
                bn = _holder(self, s.prefix + '.conv_block.1')
                bn.weight = nn.Parameter(torch.ones(s.cout))
                bn.bias = nn.Parameter(torch.zeros(s.cout))
                bn.register_buffer('running_mean', torch.zeros(s.cout))
                bn.register_buffer('running_var', torch.ones(s.cout))
                bn.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
            else:
                conv = _holder(self, s.prefix)
                conv.weight = nn.Parameter(w)
                bound = 1.0 / math.sqrt(fan_in)
                conv.bias = nn.Parameter(torch.empty(s.cout).uniform_(-bound, bound))
        for p in self.parameters():
            p.requires_grad_(False)
        # one engine (buffer plan of ~0.1 GB per image at 544x544) per distinct (batch, H, W); least recently used plans are
        # dropped beyond this many so that a stream of differently padded batches cannot grow without bound
        self.max_engines = max(1, int(os.environ.get('ORIENMASK_B200_MAX_ENGINES', '4')))
        self._engines = {}
        self._weights_version = 0
        self._fingerprint = None
        # Every forward compares a fingerprint of the live parameters / buffers with the one the packed weights were built from
        # (see _weights_fingerprint); set to False to skip the ~0.1 ms host check when the weights are known to be frozen.
        self.check_weights = True
        self.register_load_state_dict_post_hook(_invalidate_after_load)
        if pretrained is not None:
            # model/base.py:48-64: a *backbone* checkpoint (keys relative to DarkNet53: 'conv1.conv_block.0.weight', ...);
            # every key the backbone has with the same shape is taken, the others are reported and ignored
            ckpt = torch.load(pretrained, map_location='cpu')
            own = self.state_dict()
            taken = {'backbone.' + k: v for k, v in ckpt.items()
                     if 'backbone.' + k in own and tuple(v.shape) == tuple(own['backbone.' + k].shape)}
            ignored = [k for k in ckpt if 'backbone.' + k not in taken]
            own.update(taken)
            self.load_state_dict(own)
            print('[%s] Load pretrained model %s' % (type(self).__name__, pretrained))
            if ignored:
                print('Ignore keys:', ignored)

    # -- weight-version tracking: any reload or device/dtype move drops the packed weights ---------
    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._invalidate()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._invalidate()
        return r

    def _invalidate(self):
        self._engines = {}
        self._weights_version += 1
        self._fingerprint = None
        self.__dict__['_state_tensors'] = None

    def invalidate(self):
        """Drop the folded / packed weights: the next forward rebuilds them from the live parameters.  Needed only after an
        update the fingerprint cannot see (``p.data.copy_(...)``: a write through a detached alias bumps no version counter)."""
        self._invalidate()

    def _weights_fingerprint(self):
        """(version counter, storage address) of every parameter and buffer: changes on in-place updates of the tensors
        (optimizer steps, ``p.copy_``, a parent module's ``load_state_dict`` -- which never calls this module's override),
        on ``p.data = ...`` swaps and on device / dtype moves.  The reference always runs on the live parameters; the engine's
        packed copies are rebuilt whenever this differs from the fingerprint they were built from."""
        tensors = self.__dict__.get('_state_tensors')
        if tensors is None:                      # walking 270 sub-modules costs ~0.5 ms: done once per weight version (see _invalidate)
            tensors = self.__dict__['_state_tensors'] = list(self.parameters()) + list(self.buffers())
        return tuple([t._version for t in tensors] + [t.data_ptr() for t in tensors])

    def __getstate__(self):
        # copy.deepcopy / pickle (torch.save of the whole module) take the parameters, never the buffer plans: those hold
        # native plan handles and gigabytes of activation buffers, and are rebuilt on the first forward of the copy
        state = self.__dict__.copy()
        state['_engines'] = {}
        state['_state_tensors'] = None
        state['_fingerprint'] = None
        return state

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('orienmask_b200 runs on CUDA (sm_100a) only; got a %s tensor and there is no CPU fallback' % x.device)
        if x.dim() != 4 or x.size(1) != 3 or x.size(2) % 32 or x.size(3) % 32:
            raise ValueError('expected [B,3,H,W] with H, W multiples of 32, got %s' % (tuple(x.shape),))
        if self.training and not getattr(self, '_warned_training', False):
            import warnings
            warnings.warn('orienmask_b200 is an inference engine: BatchNorm uses running statistics even in train() mode '
                          '(the reference would use batch statistics); call .eval()')
            self._warned_training = True
        if self.check_weights:
            fp = self._weights_fingerprint()
            if fp != self._fingerprint:
                if self._fingerprint is not None:
                    self._engines = {}
                    self._weights_version += 1
                self._fingerprint = fp
        eng = self._engine_for((int(x.size(0)), int(x.size(2)), int(x.size(3)), self.precision, x.device.index), x.device)
        return eng.run_graph(x) if self.use_cuda_graph else eng.run(x)

    def _engine_for(self, key, device):
        eng = self._engines.pop(key, None)
        if eng is None:
            while len(self._engines) >= self.max_engines:        # dicts keep insertion order: the first key is the oldest use
                del self._engines[next(iter(self._engines))]
            # the schedule lives in the C library (om_engine_create / om_forward); ORIENMASK_B200_ENGINE=py selects the Python-scheduled
            # twin (one C-ABI call per layer) that tools and tests keep as a bit-exact cross-check
            cls = _PyEngine if os.environ.get('ORIENMASK_B200_ENGINE') == 'py' else _Engine
            eng = cls(self, *key[:3], precision=self.precision, device=device)
        self._engines[key] = eng                                  # (re-)inserted last = most recently used
        return eng


class OrienMaskYOLO(OrienMaskYOLOFPNPlus):
    """``/root/reference/model/orienmask_yolo.py:8-86``: same backbone, necks and heads; the stride-4 neck takes the up-sampled
    ``route8`` instead of the four skip convolutions (its own 506 state-dict keys)."""
    _plus = False


_PRECISIONS = {'fp16': _lib.PREC_F16, 'fp32': _lib.PREC_F32, 'parity': _lib.PREC_SPLIT}


class _Engine:
    """One (batch, H, W, precision) instance of the C library's whole-network engine (csrc/engine.cu): the state dict goes over as
    named fp32 device tensors, BN folding / weight packing / the buffer plan / the 95-launch schedule all happen behind
    om_engine_create, and a forward is ONE C-ABI call (om_forward)."""

    def __init__(self, model, B, H, W, precision, device):
        if precision not in _PRECISIONS:
            raise ValueError("precision must be 'fp16', 'parity' or 'fp32'")
        self.lib = _lib.lib()
        self.B, self.H, self.W, self.device = B, H, W, device
        self.prec = _PRECISIONS[precision]
        self.nA, self.nC, self.plus = model.num_anchors, model.num_classes, model._plus
        cfg = _lib.EngineConfig(self.prec, B, H, W, self.nA, self.nC, int(self.plus))
        sd = {k: v.detach().to(device=device, dtype=torch.float32).contiguous() for k, v in model.state_dict().items()
              if v.is_floating_point()}
        names = [k.encode() for k in sd]
        arr = (_lib.OmTensor * len(sd))(*[_lib.OmTensor(n, t.data_ptr(), t.numel()) for n, t in zip(names, sd.values())])
        nbytes = ctypes.c_size_t(0)
        _lib.check(self.lib.om_engine_workspace_bytes(ctypes.byref(cfg), ctypes.byref(nbytes)), 'om_engine_workspace_bytes')
        self.handle = _lib.c_vp()
        with torch.cuda.device(device):
            self.workspace = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
            _lib.check(self.lib.om_engine_create(ctypes.byref(cfg), arr, len(sd), _lib.ptr(self.workspace), nbytes.value,
                                                 _lib.stream_ptr(), ctypes.byref(self.handle)), 'om_engine_create')
            torch.cuda.current_stream(device).synchronize()       # the fold / pack kernels have read the state dict
        nb = self.nA * (5 + self.nC)
        self.out_shapes = [(B, nb, H // s, W // s) for s in (32, 16, 8)] + [(B, self.nA * 6, H // 4, W // 4)]
        self._static = None
        self._layers = None

    @property
    def layers(self):
        """One dict per launch, in launch order (tools/layer_report.py): name, shape, flops, bytes."""
        if self._layers is None:
            out = []
            info = _lib.LayerInfo()
            for i in range(self.lib.om_engine_layer_count(self.handle)):
                _lib.check(self.lib.om_engine_layer_info(self.handle, i, ctypes.byref(info)), 'om_engine_layer_info')
                out.append(dict(name=info.name.decode(), shape=info.shape.decode(), flops=info.flops, bytes=info.bytes,
                                conv=info.conv, is_stem=bool(info.is_stem)))
            self._layers = out
        return self._layers

    @property
    def flops(self):
        return sum(l['flops'] for l in self.layers)

    def _outputs(self):
        return [torch.empty(s, dtype=torch.float32, device=self.device) for s in self.out_shapes]

    def _tuple(self, outs):
        n2 = self.nA * 2
        o = outs[3]
        return ((outs[0], o[:, 0:n2]), (outs[1], o[:, n2:2 * n2]), (outs[2], o[:, 2 * n2:3 * n2]))

    def run(self, x, fresh=True):
        """fresh=True: the head tensors are allocated for this call (the caller owns them, like the reference's forward);
        fresh=False (CUDA-graph capture): the engine's static tensors, overwritten by the next call."""
        x = x.contiguous().float()
        if fresh:
            outs = self._outputs()
        else:
            if self._static is None:
                self._static = self._outputs()
            outs = self._static
        with torch.cuda.device(self.device):
            bbox = (_lib.c_vp * 3)(*[t.data_ptr() for t in outs[:3]])
            _lib.check(self.lib.om_forward(self.handle, _lib.ptr(x), bbox, _lib.ptr(outs[3]), _lib.stream_ptr()), 'om_forward')
        return self._tuple(outs)

    def run_graph(self, x):
        """Same as run(), replayed from a CUDA graph captured on first use (input copied into a static buffer)."""
        if getattr(self, '_graph', None) is None:
            self._static_x = torch.empty(self.B, 3, self.H, self.W, dtype=torch.float32, device=self.device)
            self._static_x.copy_(x)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.run(self._static_x, fresh=False)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_out = self.run(self._static_x, fresh=False)
            self._graph = g
        self._static_x.copy_(x)
        self._graph.replay()
        return self._graph_out

    def time_layers(self, x, iters=5):
        """Per-launch device times (us, mean over `iters` passes) from CUDA events between the launches."""
        n = self.lib.om_engine_layer_count(self.handle)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(iters)]
        x = x.contiguous().float()
        outs = self._outputs()
        with torch.cuda.device(self.device):
            stream = _lib.stream_ptr()
            bbox = (_lib.c_vp * 3)(*[t.data_ptr() for t in outs[:3]])
            for it in range(iters):
                ev[it][0].record()
                for i in range(n):
                    _lib.check(self.lib.om_engine_run_layer(self.handle, i, _lib.ptr(x), bbox, _lib.ptr(outs[3]), stream), 'om_engine_run_layer')
                    ev[it][i + 1].record()
            torch.cuda.synchronize()
        return [1e3 * sum(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(iters)) / iters for i in range(n)]

    def __del__(self):
        try:
            if self.handle:
                self.lib.om_engine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _PyEngine:
    """The same schedule driven from Python, one C-ABI call per layer (om_conv_create / om_conv_run): the engine of round 1, kept as
    the bit-exact twin of csrc/engine.cu (tests/test_gpu_forward.py::test_c_engine_matches_the_python_schedule) and as the schedule the
    GPU-less planner sweep walks (tools/plan_table.py)."""

    def __init__(self, model, B, H, W, precision, device):
        if precision not in ('fp16', 'fp32', 'parity'):
            raise ValueError("precision must be 'fp16', 'parity' or 'fp32'")
        self.lib = _lib.lib()
        self.B, self.H, self.W, self.device = B, H, W, device
        self.prec = {'fp16': _lib.PREC_F16, 'fp32': _lib.PREC_F32, 'parity': _lib.PREC_SPLIT}[precision]
        self.adt = torch.float32 if precision == 'fp32' else torch.float16
        self.cmul = 2 if precision == 'parity' else 1         # halves per activation value (split precision: hi | lo)
        self.nA, self.nC = model.num_anchors, model.num_classes
        self.plus = model._plus
        self.sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in model.state_dict().items()
                   if v.is_floating_point()}
        self.keep = []            # owns every device tensor the plans point to
        self.plans = []           # ('stem', args) | ('conv', handle)
        self.head_slot = {}       # plan index of a head layer -> output slot (0..2 bbox scales, 3 orientation)
        self.flops = 0
        self.layers = []          # one entry per launch, in launch order (tools/layer_report.py)
        with torch.cuda.device(device):
            self._build()
        self.sd = None

    # ---- buffers ---------------------------------------------------------------------------------
    def rows(self, stride):
        """Rows per image of the padded-row layout at this stride.  Strides 1 / 2 / 4 are tied by the parity-split stride-2
        layers (rows(in) == 2 * rows(out)); the stride-4 maps carry ONE pad row (the 3x3 layers there hold a third of all FLOPs
        and every pad row is computed), which fixes 4 and 2 pad rows at strides 2 and 1.  Strides 8 / 16 / 32 keep
        rows(s) == 2 * rows(2s) so that their up-add sources can be staged by TMA."""
        tight = os.environ.get('ORIENMASK_B200_TIGHT_ROWS', '1')
        if stride <= 4 and tight in ('1', '2'):
            return (self.H // 4 + 1) * (4 // stride)
        if stride == 8 and tight == '2':
            return self.H // 8 + 1
        return self.H // stride + 32 // stride

    def act(self, stride, channels, dtype=None, s2d=False):
        """Padded-row NHWC buffer; s2d=True marks it parity-split (same size: four [B*rows/2, W/2, C] planes)."""
        t = torch.zeros(self.B * self.rows(stride), self.W // stride, channels * (self.cmul if dtype is None else 1),
                        dtype=dtype or self.adt, device=self.device)
        self.keep.append(t)
        return dict(t=t, stride=stride, c=channels, s2d=s2d)

    # ---- weights ---------------------------------------------------------------------------------
    def folded(self, prefix, kind):
        sd = self.sd
        if kind == 'cbl':
            w = sd[prefix + '.conv_block.0.weight']
            g, b = sd[prefix + '.conv_block.1.weight'], sd[prefix + '.conv_block.1.bias']
            mu, var = sd[prefix + '.conv_block.1.running_mean'], sd[prefix + '.conv_block.1.running_var']
            scale = g / torch.sqrt(var + 1e-5)
            return w * scale.view(-1, 1, 1, 1), b - mu * scale
        return sd[prefix + '.weight'], sd[prefix + '.bias']

    def pack(self, w):
        """[cout, cin, k, k] fp32 -> (engine layout, accumulator scale); see _lib.pack_conv_weights / om_conv_desc.weights."""
        p, acc_scale = _lib.pack_conv_weights(w.to(self.device), self.prec)
        self.keep.append(p)
        return p, acc_scale

    # ---- op emission -----------------------------------------------------------------------------
    def conv(self, src, w, bias, dst, k, stride=1, leaky=True, kind=_lib.OUT_ACT, residual=None, upadd=None, nchw=None, name=None):
        d = _lib.ConvDesc()
        d.precision, d.batch = self.prec, self.B
        si = src['stride']
        so = si * stride
        d.in_h, d.in_w, d.in_rows = self.H // si, self.W // si, self.rows(si)
        d.out_h, d.out_w, d.out_rows = self.H // so, self.W // so, self.rows(so)
        d.cin, d.cout = w.shape[1], w.shape[0]
        d.ksize, d.stride, d.leaky, d.out_kind = k, stride, int(leaky), kind
        d.in_s2d = int(bool(src.get('s2d')))
        d.out_s2d = int(bool(dst is not None and dst.get('s2d')))
        assert not d.in_s2d or stride == 2, 'only a stride-2 layer can read a parity-split buffer'
        d.input = src['t'].data_ptr()
        wp, d.acc_scale = self.pack(w)
        d.weights = wp.data_ptr()
        if bias is not None:
            b = bias.contiguous().clone()
            self.keep.append(b)
            d.bias = b.data_ptr()
        if kind == _lib.OUT_NCHW:
            d.cout_stride = d.cout
            d.output = nchw.data_ptr()
        else:
            assert dst['stride'] == so and dst['c'] == d.cout
            d.cout_stride = dst['c']
            d.output = dst['t'].data_ptr()
        if residual is not None:
            d.residual = residual['t'].data_ptr()
        if upadd is not None:
            assert upadd['stride'] == 2 * so and upadd['c'] == d.cout
            d.upadd = upadd['t'].data_ptr()
            d.up_rows = self.rows(upadd['stride'])
        handle = _lib.c_vp()
        _lib.check(self.lib.om_conv_create(d, handle), 'om_conv_create')
        self.plans.append(('conv', handle))
        flops = 2 * d.batch * d.out_h * d.out_w * d.cout * d.cin * k * k
        self.flops += flops
        esz = 2 if self.prec == _lib.PREC_F16 else 4           # split precision: two halves per value
        nbytes = d.batch * d.in_h * d.in_w * d.cin * esz + w.numel() * esz
        nbytes += d.batch * d.out_h * d.out_w * d.cout * (esz if kind == _lib.OUT_ACT else 4)
        if residual is not None:
            nbytes += d.batch * d.out_h * d.out_w * d.cout * esz
        if upadd is not None:
            nbytes += d.batch * d.out_h * d.out_w * d.cout        # fp32 quarter-resolution partial
        self.layers.append(dict(name=name or '?', shape='%dx%d s%d %d->%d @%dx%d%s%s%s' % (
            k, k, stride, d.cin, d.cout, d.out_h, d.out_w, ' +res' if residual is not None else '',
            ' +up' if upadd is not None else '', ('', ' partial', ' nchw')[kind]), flops=flops, bytes=nbytes))

    def cbl(self, prefix, src, dst, k, stride=1, residual=None):
        w, b = self.folded(prefix, 'cbl')
        self.conv(src, w, b, dst, k, stride, True, residual=residual, name=prefix)

    def chain(self, prefix, src, bufs, ks, first_upadd=None, first_cols=None):
        """conv_bn_leaky sequence <prefix>.0.. ; bufs alternate; optional concat-split on the first conv."""
        cur = src
        for i, k in enumerate(ks):
            w, b = self.folded('%s.%d' % (prefix, i), 'cbl')
            dst = bufs[i % 2]
            if i == 0 and first_cols is not None:
                w = w[:, first_cols[0]:first_cols[1]].contiguous()
            self.conv(cur, w, b, dst, k, upadd=first_upadd if i == 0 else None, name='%s.%d' % (prefix, i))
            cur = dst
        return cur

    def partial(self, prefix, cols, src, upadd=None):
        """fp32 pre-activation partial W[:, cols] * src (+ nearest-up of a coarser partial)."""
        w, _ = self.folded(prefix, 'cbl')
        w = w[:, cols[0]:cols[1]].contiguous()
        dst = self.act(src['stride'], w.shape[0], torch.float32)
        self.conv(src, w, None, dst, 1, leaky=False, kind=_lib.OUT_PARTIAL, upadd=upadd, name='%s[%d:%d]' % (prefix, cols[0], cols[1]))
        return dst

    def head(self, prefix, src, out, slot):
        w, b = self.folded(prefix, 'conv')
        self.head_slot[len(self.plans)] = slot
        self.conv(src, w, b, None, 1, leaky=False, kind=_lib.OUT_NCHW, nchw=out, name=prefix)

    # ---- the schedule (model/orienmask_yolo_fpnplus.py:74-90, model/backbone/darknet.py:47-54) ----
    def _build(self):
        B, H, W, dev = self.B, self.H, self.W, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        w1, b1 = self.folded('backbone.conv1', 'cbl')
        self.stem_w = w1.permute(2, 3, 1, 0).reshape(27, 32).contiguous()
        self.stem_b = b1.contiguous()
        # The outputs of the stem and of stage conv2 are read only by the next stage's stride-2 convolution: they are
        # written parity-split so that each of its taps is a dense TMA box (strided gathers ran at ~18 B/clk/SM).
        c1 = self.act(1, 32, s2d=True)
        self.plans.append(('stem', c1))
        self.flops += 2 * B * H * W * 32 * 27
        self.layers.append(dict(name='backbone.conv1', shape='3x3 s1 3->32 @%dx%d stem' % (H, W), flops=2 * B * H * W * 32 * 27,
                                bytes=B * H * W * (3 * 4 + 32 * (2 if self.prec == _lib.PREC_F16 else 4))))

        trunk = c1
        feats = {}
        for i, (c, n) in enumerate(zip(STAGE_CHANNELS, STAGE_BLOCKS)):
            stage = 'backbone.conv%d' % (i + 2)
            st = 2 ** (i + 1)
            x = self.act(st, 2 * c)
            y = self.act(st, c)
            self.cbl(stage + '.0', trunk, x, 3, stride=2)
            for b in range(1, n + 1):
                self.cbl('%s.%d.conv.0' % (stage, b), x, y, 1)
                if i == 0 and b == n:                                            # feeds only conv3.0 (stride 2)
                    xs = self.act(st, 2 * c, s2d=True)
                    self.cbl('%s.%d.conv.1' % (stage, b), y, xs, 3, residual=x)
                    x = xs
                else:
                    self.cbl('%s.%d.conv.1' % (stage, b), y, x, 3, residual=x)  # in place: x += leaky(conv(y))
            trunk = x
            feats[st] = x
        x4, x8, x16, x32 = feats[4], feats[8], feats[16], feats[32]
        ks = (1, 3, 1, 3, 1)

        neck32 = self.chain('neck32', x32, (self.act(32, 512), self.act(32, 1024)), ks)
        r32 = self.act(32, 256)
        self.cbl('route32.0', neck32, r32, 1)
        p16 = self.partial('neck16.0', (0, 256), r32)
        neck16 = self.chain('neck16', x16, (self.act(16, 256), self.act(16, 512)), ks, first_upadd=p16, first_cols=(256, 768))
        r16 = self.act(16, 128)
        self.cbl('route16.0', neck16, r16, 1)
        p8 = self.partial('neck8.0', (0, 128), r16)
        neck8 = self.chain('neck8', x8, (self.act(8, 128), self.act(8, 256)), ks, first_upadd=p8, first_cols=(128, 384))

        nb = self.nA * (5 + self.nC)
        self.out_bbox = []
        for st, neck, c in ((32, neck32, 512), (16, neck16, 256), (8, neck8, 128)):
            hb = self.act(st, 2 * c)
            self.cbl('bbox_head%d.0' % st, neck, hb, 3)
            out = torch.empty(B, nb, H // st, W // st, **f32)
            self.head('bbox_head%d.1' % st, hb, out, len(self.out_bbox))
            self.out_bbox.append(out)

        a4, b4 = self.act(4, 128), self.act(4, 256)
        if self.plus:
            s32, s16, s8, s4 = self.act(32, 64), self.act(16, 64), self.act(8, 64), self.act(4, 64)
            self.cbl('skip32.0', neck32, s32, 1)
            self.cbl('skip16.0', neck16, s16, 1)
            self.cbl('skip8.0', neck8, s8, 1)
            self.cbl('skip4', x4, s4, 1)
            q32 = self.partial('neck4.0', (0, 64), s32)
            q16 = self.partial('neck4.0', (64, 128), s16, upadd=q32)
            q8 = self.partial('neck4.0', (128, 192), s8, upadd=q16)
            neck4 = self.chain('neck4', s4, (a4, b4), ks, first_upadd=q8, first_cols=(192, 256))
        else:                         # model/orienmask_yolo.py:83: neck4(cat[route8(neck8) up2, x4])
            r8 = self.act(8, 64)
            self.cbl('route8.0', neck8, r8, 1)
            p4 = self.partial('neck4.0', (0, 64), r8)
            neck4 = self.chain('neck4', x4, (a4, b4), ks, first_upadd=p4, first_cols=(64, 192))
        # orien_head.0-4 alternate 3x3 (128->256) and 1x1 (256->128): neck4 lives in a4, so start on b4
        o = self.chain('orien_head', neck4, (b4, a4), (3, 1, 3, 1, 3))
        self.out_orien = torch.empty(B, self.nA * 6, H // 4, W // 4, **f32)
        self.head('orien_head.5', o, self.out_orien, 3)

    def run_graph(self, x):
        """Same as run(), replayed from a CUDA graph captured on first use (input copied into a static buffer)."""
        if getattr(self, '_graph', None) is None:
            self._static_x = torch.empty(self.B, 3, self.H, self.W, dtype=torch.float32, device=self.device)
            self._static_x.copy_(x)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.run(self._static_x, fresh=False)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_out = self.run(self._static_x, fresh=False)
            self._graph = g
        self._static_x.copy_(x)
        self._graph.replay()
        return self._graph_out

    def run(self, x, fresh=True):
        """fresh=True: the head tensors are allocated for this call (the caller owns them, like the reference's forward);
        fresh=False (CUDA-graph capture): the engine's static tensors, overwritten by the next call."""
        x = x.contiguous().float()
        with torch.cuda.device(self.device):
            stream = _lib.stream_ptr()
            outs = ([torch.empty_like(t) for t in self.out_bbox] + [torch.empty_like(self.out_orien)]) if fresh else None
            for i, (kind, arg) in enumerate(self.plans):
                if kind == 'conv' and outs is not None and i in self.head_slot:
                    _lib.check(self.lib.om_conv_run_to(arg, _lib.ptr(outs[self.head_slot[i]]), stream), 'om_conv_run_to')
                elif kind == 'conv':
                    _lib.check(self.lib.om_conv_run(arg, stream), 'om_conv_run')
                else:
                    _lib.check(self.lib.om_stem_conv(self.prec, _lib.ptr(x), _lib.ptr(self.stem_w), _lib.ptr(self.stem_b),
                                                     _lib.ptr(arg['t']), self.B, self.H, self.W, self.rows(1), 32, int(bool(arg.get('s2d'))), stream),
                               'om_stem_conv')
        n2 = self.nA * 2
        bbox, o = (outs[:3], outs[3]) if outs is not None else (self.out_bbox, self.out_orien)
        return ((bbox[0], o[:, 0:n2]), (bbox[1], o[:, n2:2 * n2]), (bbox[2], o[:, 2 * n2:3 * n2]))

    def time_layers(self, x, iters=5):
        """Per-launch device times (us, mean over `iters` passes) from CUDA events between the launches."""
        n = len(self.plans)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(iters)]
        x = x.contiguous().float()
        with torch.cuda.device(self.device):
            stream = _lib.stream_ptr()
            for it in range(iters):
                ev[it][0].record()
                for i, (kind, arg) in enumerate(self.plans):
                    if kind == 'conv':
                        _lib.check(self.lib.om_conv_run(arg, stream), 'om_conv_run')
                    else:
                        _lib.check(self.lib.om_stem_conv(self.prec, _lib.ptr(x), _lib.ptr(self.stem_w), _lib.ptr(self.stem_b),
                                                         _lib.ptr(arg['t']), self.B, self.H, self.W, self.rows(1), 32, int(bool(arg.get('s2d'))), stream),
                                   'om_stem_conv')
                    ev[it][i + 1].record()
            torch.cuda.synchronize()
        return [1e3 * sum(ev[it][i].elapsed_time(ev[it][i + 1]) for it in range(iters)) / iters for i in range(n)]

    def __del__(self):
        try:
            for kind, arg in self.plans:
                if kind == 'conv':
                    self.lib.om_conv_destroy(arg)
        except Exception:
            pass
