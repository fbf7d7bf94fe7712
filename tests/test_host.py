"""CPU-side tests: the C-ABI library loads and exports every declared symbol, the host mirrors keep the
reference interface, sharding logic (gloo, world size 2)."""
import ctypes
import types
import functools
import os
import re
import sys

import pytest
import torch

from tests.common import ROOT, post_config


def test_library_exports_every_declared_symbol():
    from orienmask_b200 import _lib, build
    build.build()
    header = open(os.path.join(ROOT, 'include', 'orienmask_b200.h')).read()
    declared = set(re.findall(r'\b(om_[a-z0-9_]+)\s*\(', header))
    declared -= {'om_conv'}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().om_abi_version() == 8


def test_header_is_plain_c_and_links(tmp_path):
    """include/orienmask_b200.h compiles as strict C99 in a caller without CUDA or torch, every declared entry point
    resolves at link time, and the struct layouts the C compiler sees are the ones the ctypes binding declares."""
    import subprocess
    from orienmask_b200 import _lib, build
    build.build()
    exe = str(tmp_path / 'abi_check')
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-Werror', '-pedantic',
                           '-I', os.path.join(ROOT, 'include'), os.path.join(ROOT, 'tests', 'c_abi', 'abi_check.c'),
                           '-o', exe, '-L', libdir, '-lorienmask_b200', '-Wl,-rpath,' + libdir])
    words = subprocess.check_output([exe]).decode().split()
    got = {words[i]: int(words[i + 1]) for i in range(0, len(words), 2)}
    assert got["abi"] == 8 and got['entries'] == len(_lib.SIGNATURES)
    assert got['sizeof(om_post_config)'] == ctypes.sizeof(_lib.PostConfig)
    assert got['sizeof(om_conv_desc)'] == ctypes.sizeof(_lib.ConvDesc)
    assert got['sizeof(om_prep_config)'] == ctypes.sizeof(_lib.PrepConfig)
    assert got['sizeof(om_rle_image)'] == ctypes.sizeof(_lib.RleImage)
    assert got['sizeof(om_blend_config)'] == ctypes.sizeof(_lib.BlendConfig)
    for name, (struct, field) in {'om_post_config.anchor_w': (_lib.PostConfig, 'anchor_w'), 'om_post_config.nms_post': (_lib.PostConfig, 'nms_post'),
                                  'om_conv_desc.input': (_lib.ConvDesc, 'input'), 'om_conv_desc.out_s2d': (_lib.ConvDesc, 'out_s2d'),
                                  'om_prep_config.pad_value': (_lib.PrepConfig, 'pad_value'), 'om_rle_image.vflip': (_lib.RleImage, 'vflip'),
                                  'om_blend_config.alpha': (_lib.BlendConfig, 'alpha')}.items():
        assert got['offsetof(%s)' % name] == getattr(struct, field).offset, name


def test_fastdiv_is_exact(tmp_path):
    """csrc/common.cuh FastDiv (tile decode by multiply-high in every persistent kernel): n / d for all 0 <= n < 2^31 -- checked here
    on the host for every divisor up to 70 000 (tile counts, row pitches, channel chunks) against corner and random numerators."""
    import subprocess
    exe = str(tmp_path / 'fastdiv_check')
    cuda_inc = '/usr/local/cuda/include'
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-I', cuda_inc, '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'c_abi', 'fastdiv_check.cpp'), '-o', exe])
    assert subprocess.check_output([exe]).decode().strip() == 'bad 0'


def test_built_library_contains_blackwell_tensor_and_tma_code():
    """SASS of the in-tree .so (cuobjdump, no GPU needed): the conv engine is tcgen05 (UTCHMMA.2CTA) fed by TMA (UTMALDG) with
    accumulators read from TMEM (LDTM); no kernel uses the legacy mma.sync path (HMMA) -- B200_PROFILING.md's mnemonics."""
    import shutil
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    try:
        import sass_report
    finally:
        sys.path.pop(0)
    from orienmask_b200 import build
    build.build()
    counts = dict(sass_report.sass_counts())
    assert len(counts) >= 20
    for name in ('conv_tc2_kernel<32, false>', 'conv_tc2_kernel<64, false>', 'conv_tc2_kernel<32, true>', 'conv_tc2_kernel<64, true>'):
        c = counts[name]
        assert c['UTCHMMA'] > 100 and c['UTMALDG'] >= 5 and c['LDTM'] >= 4 and c['UTCBAR'] >= 2 and c['ACQBULK'] >= 1, (name, dict(c))
    assert counts['stem_tc_kernel<true, false>']['UTCHMMA'] >= 2 and counts['stem_tc_kernel<true, false>']['LDTM'] >= 1
    assert counts['stem_tc_kernel<true, false>']['UTMALDG'] >= 1            # the 544x544 input tiles are staged by TMA
    for name in ('stem_fused_kernel', 'dark_block_kernel'):                 # the two fused launches: two MMA stages each, TMA-fed, FADD2 / FMUL2 epilogues
        c = counts[name]
        assert c['UTCHMMA'] >= 20 and c['UTMALDG'] >= 1 and c['LDTM'] >= 2, (name, dict(c))
    assert all(c['HMMA'] == 0 and c['HGMMA'] == 0 for c in counts.values())
    for name in ('mask_kernel', 'conf_compact_kernel', 'select_edge_kernel', 'select_tail_kernel', 'nms_kernel<true>', 'prep_kernel<unsigned char>',
                 'mask_rle_kernel', 'mask_blend_kernel'):
        assert name in counts, name


def test_config_errors_are_reported_without_a_gpu():
    from orienmask_b200 import _lib
    lib = _lib.lib()
    cfg = _lib.PostConfig()
    n = ctypes.c_size_t(0)
    rc = lib.om_post_workspace_bytes(ctypes.byref(cfg), 1, ctypes.byref(n))
    assert rc == -1 and b'num_scales' in lib.om_last_error()


def test_conv_descriptor_errors_are_reported_without_a_gpu():
    """om_conv_create validates the descriptor before it touches CUDA: each bad field gives OM_ERR_* and a message."""
    from orienmask_b200 import _lib
    lib = _lib.lib()

    def desc(**kw):
        d = _lib.ConvDesc()
        d.precision, d.batch = _lib.PREC_F16, 1
        d.in_h = d.in_w = d.out_h = d.out_w = 16
        d.in_rows = d.out_rows = 17
        d.cin, d.cout, d.cout_stride, d.ksize, d.stride, d.leaky, d.out_kind = 64, 64, 64, 3, 1, 1, _lib.OUT_ACT
        d.input = d.weights = d.output = 256
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    cases = [(dict(ksize=5), -3, b'ksize'), (dict(stride=3), -3, b'stride'), (dict(precision=7), -1, b'precision'),
             (dict(in_rows=16), -1, b'in_rows'), (dict(out_h=8), -1, b'geometry'), (dict(input=0), -1, b'null tensor'),
             (dict(cout_stride=32), -1, b'cout_stride'), (dict(cin=48), -1, b'cin'), (dict(batch=0), -1, b'non-positive'),
             (dict(residual=256, out_kind=_lib.OUT_NCHW), -1, b'residual'), (dict(in_s2d=1), -1, b'in_s2d'),
             # split precision stores hi | lo halves of a dense pixel: a channel pitch wider than cout is refused
             (dict(precision=_lib.PREC_SPLIT, cout_stride=128), -1, b'split precision needs a dense activation output')]
    for kw, code, text in cases:
        handle = _lib.c_vp()
        rc = lib.om_conv_create(desc(**kw), handle)
        assert rc == code and text in lib.om_last_error(), (kw, rc, lib.om_last_error())
        assert not handle.value
    assert lib.om_conv_run(None, None) == -1 and lib.om_conv_run_to(None, None, None) == -1
    lib.om_conv_destroy(None)


def test_entry_point_argument_errors_without_a_gpu():
    """Every C-ABI entry point rejects bad arguments with OM_ERR_* + a message before any CUDA call."""
    import orienmask_b200 as ob
    from orienmask_b200 import _lib
    lib = _lib.lib()
    post = ob.OrienMaskYOLOPostProcess(**post_config(544, 544))
    cfg, null = ctypes.byref(post._cfg), None

    def err(rc, text, code=-1):
        assert rc == code and text in lib.om_last_error(), (rc, lib.om_last_error())

    n = ctypes.c_size_t(0)
    assert lib.om_post_workspace_bytes(cfg, 32, ctypes.byref(n)) == 0
    assert n.value >= 32 * 18207 * 80 * 16                         # 16 bytes per (prediction, class) pair and image
    err(lib.om_post_workspace_bytes(cfg, 0, ctypes.byref(n)), b'bad batch')
    err(lib.om_decode_select(cfg, null, null, 1, null, null, null, null, null, null), b'om_decode_select')
    err(lib.om_batched_nms(cfg, null, null, null, null, 1, null, null, null, null, null, null, null), b'om_batched_nms')
    err(lib.om_mask_assemble(cfg, null, null, null, null, null, 1, null, null), b'om_mask_assemble')
    err(lib.om_nms(null, 2000, 0.5, null, null, null), b'outside [0, 1024]')
    err(lib.om_nms(null, 4, 0.5, null, null, null), b'null argument')
    bad = _lib.PostConfig.from_buffer_copy(post._cfg)
    bad.nms_post = 500
    err(lib.om_post_workspace_bytes(ctypes.byref(bad), 1, ctypes.byref(n)), b'nms_post')
    bad = _lib.PostConfig.from_buffer_copy(post._cfg)
    bad.image_h = 500
    err(lib.om_post_workspace_bytes(ctypes.byref(bad), 1, ctypes.byref(n)), b'multiple of 32')
    prep = _lib.PrepConfig()
    err(lib.om_preprocess(ctypes.byref(prep), null, 0, 1, null, null), b'om_preprocess')
    prep.src_h = prep.src_w = prep.resize_h = prep.resize_w = 8
    prep.out_h = prep.out_w = 4
    prep.src_dtype = _lib.SRC_U8
    err(lib.om_preprocess(ctypes.byref(prep), 256, 0, 1, 256, null), b'does not fit')
    err(lib.om_mask_rle(null, 1, 1, 1, 1, 1, 1, 1, null, null, null, null, null), b'om_mask_rle')
    blend = _lib.BlendConfig()
    err(lib.om_mask_areas(ctypes.byref(blend), null, 1, null, null, null), b'')
    err(lib.om_mask_blend(ctypes.byref(blend), null, 1, null, null, null, null), b'')
    err(lib.om_stem_conv(_lib.PREC_F16, 256, 256, 256, 256, 1, 16, 32, 17, 64, 0, null), b'cout must be 32', code=-3)
    err(lib.om_stem_conv(_lib.PREC_F16, 256, 256, 256, 256, 1, 16, 32, 16, 32, 0, null), b'bad geometry')


def test_model_state_dict_layout_and_loud_cpu_failure():
    import orienmask_b200 as ob
    from orienmask_b200.arch import state_dict_shapes, macs_per_image
    from orienmask_b200.synthetic import synthetic_state_dict
    m = ob.OrienMaskYOLOFPNPlus(num_anchors=3, num_classes=80, pretrained=None, freeze_backbone=False,
                                backbone_batchnorm_eval=False)
    sd = m.state_dict()
    shapes = state_dict_shapes()
    assert len(sd) == 524 and set(sd) == set(shapes)
    m.load_state_dict(synthetic_state_dict(0), strict=True)
    assert abs(macs_per_image(544, 544)[0] - 86.9226e9) < 1e6          # 86.9226 GMAC
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.zeros(1, 3, 64, 64))


def test_pretrained_backbone_checkpoint_is_loaded_like_the_reference(tmp_path, capsys):
    """model/base.py:48-64 + darknet.py:39: `pretrained` names a DarkNet-53 checkpoint whose keys are relative to the backbone;
    matching keys are taken, shape mismatches and unknown keys are ignored (and listed)."""
    import orienmask_b200 as ob
    ckpt = {'conv1.conv_block.0.weight': torch.full((32, 3, 3, 3), 0.25), 'conv1.conv_block.1.running_var': torch.full((32,), 3.0),
            'conv2.0.conv_block.0.weight': torch.zeros(1, 1), 'fc.weight': torch.zeros(1000, 1024)}
    path = str(tmp_path / 'darknet53.pth')
    torch.save(ckpt, path)
    m = ob.OrienMaskYOLOFPNPlus(3, 80, pretrained=path)
    sd = m.state_dict()
    assert torch.equal(sd['backbone.conv1.conv_block.0.weight'], ckpt['conv1.conv_block.0.weight'])
    assert torch.equal(sd['backbone.conv1.conv_block.1.running_var'], ckpt['conv1.conv_block.1.running_var'])
    assert sd['backbone.conv2.0.conv_block.0.weight'].shape == (64, 32, 3, 3)
    out = capsys.readouterr().out
    assert 'Ignore keys' in out and 'fc.weight' in out and 'conv2.0.conv_block.0.weight' in out


def test_model_copies_and_pickles_without_its_buffer_plans():
    import copy
    import pickle
    import orienmask_b200 as ob
    m = ob.OrienMaskYOLOFPNPlus(3, 80)
    m._engines[(1, 64, 64, 'fp16', 0)] = ctypes.c_void_p(1234)          # a native handle: neither copyable state nor picklable
    twin = copy.deepcopy(m)
    assert twin._engines == {} and len(m._engines) == 1
    assert torch.equal(twin.state_dict()['backbone.conv1.conv_block.0.weight'], m.state_dict()['backbone.conv1.conv_block.0.weight'])
    back = pickle.loads(pickle.dumps(m))
    assert back._engines == {} and len(back.state_dict()) == 524 and back.num_classes == 80


def test_engine_plans_are_bounded_lru(monkeypatch):
    """A stream of differently shaped batches keeps at most max_engines buffer plans, dropping the least recently used."""
    import orienmask_b200 as ob
    from orienmask_b200 import model as mod
    built = []

    class FakeEngine:
        def __init__(self, model, B, H, W, precision, device):
            built.append((B, H, W))
    monkeypatch.setattr(mod, '_Engine', FakeEngine)
    monkeypatch.setenv('ORIENMASK_B200_MAX_ENGINES', '2')
    m = ob.OrienMaskYOLOFPNPlus(3, 80)
    key = lambda b: (b, 64, 64, 'fp16', 0)
    a = m._engine_for(key(1), 'cuda:0')
    m._engine_for(key(2), 'cuda:0')
    assert m._engine_for(key(1), 'cuda:0') is a and len(built) == 2        # hit: nothing rebuilt, (1) becomes most recent
    m._engine_for(key(3), 'cuda:0')                                          # evicts (2), the least recently used
    assert list(m._engines) == [key(1), key(3)]
    assert m._engine_for(key(1), 'cuda:0') is a and len(built) == 3
    m.load_state_dict(m.state_dict())                                        # new weights: every packed plan is dropped
    assert not m._engines


def test_weight_updates_drop_the_packed_weights(monkeypatch):
    """The reference nn.Module always runs on its live parameters.  The engine's folded / packed copies must follow every update
    PyTorch can make without calling this module's own load_state_dict: a parent's load_state_dict, in-place writes (optimizer steps,
    EMA swaps through p.copy_), `p.data = ...` swaps.  Each forward compares a (version, address) fingerprint (model.py)."""
    import torch.nn as nn
    import orienmask_b200 as ob
    from orienmask_b200 import model as mod
    built = []

    class FakeEngine:
        def __init__(self, model, B, H, W, precision, device):
            built.append(model._weights_version)

        def run(self, x):
            return 'ran'
    monkeypatch.setattr(mod, '_Engine', FakeEngine)
    m = ob.OrienMaskYOLOFPNPlus(3, 80).eval()
    x = types.SimpleNamespace(is_cuda=True, dim=lambda: 4, size=lambda i: (1, 3, 64, 64)[i], device=types.SimpleNamespace(index=0))
    assert m(x) == 'ran' and m(x) == 'ran' and len(built) == 1               # unchanged weights: the plan is reused
    w = m.state_dict()['backbone.conv1.conv_block.0.weight']
    with torch.no_grad():
        w.mul_(2.0)                                                           # in place (what an optimizer step / EMA copy_ does)
    m(x)
    assert len(built) == 2
    parent = nn.Sequential(m)
    parent.load_state_dict(parent.state_dict())                               # recurses through _load_from_state_dict, not m.load_state_dict
    m(x)
    assert len(built) == 3
    p = next(m.parameters())
    p.data = p.data.clone()                                                   # storage swap: same version counter, new address
    m(x)
    assert len(built) == 4
    m(x)
    assert len(built) == 4
    with torch.no_grad():
        p.data.copy_(p.data * 3)                                              # through a detached alias: no counter, no new address ...
    m(x)
    assert len(built) == 4
    m.invalidate()                                                            # ... the documented escape hatch
    m(x)
    assert len(built) == 5
    m.train()
    with pytest.warns(UserWarning, match='inference engine'):
        m(x)


def test_postprocess_constructor_mirrors_reference():
    import orienmask_b200 as ob
    cfg = post_config(544, 544)
    p = ob.OrienMaskYOLOPostProcess(nms_func=functools.partial(ob.batched_nms, threshold=0.45), device=None, **cfg)
    assert abs(p.nms_thresh - 0.45) < 1e-9 and p.nms_pre == 400 and p.nms_post == 100
    assert ob.OrienMaskYOLOPostProcess(**cfg).nms_thresh == 0.5
    with pytest.raises(NotImplementedError):
        ob.OrienMaskYOLOPostProcess(nms_func=lambda d, c: (d, c, None), **cfg)


def test_dropin_packages_resolve_reference_names():
    """trainer/builder.py:61-77 does getattr(model, 'OrienMaskYOLOFPNPlus'), getattr(eval, 'OrienMaskYOLOPostProcess'),
    getattr(eval.function, 'batched_nms')."""
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); import model, eval, eval.function as f; from eval.coco_eval import COCOMetrics; "
            "print(model.OrienMaskYOLOFPNPlus.__module__, eval.OrienMaskYOLOPostProcess.__module__, f.batched_nms.__module__, "
            "model.OrienMaskYOLO.__module__, COCOMetrics.__module__)"
            % os.path.join(ROOT, 'orienmask_b200', 'dropin'))
    out = subprocess.check_output([sys.executable, '-c', code], cwd='/tmp').decode().split()
    assert out == ['orienmask_b200.model', 'orienmask_b200.postprocess', 'orienmask_b200.function', 'orienmask_b200.model',
                   'orienmask_b200.coco_format']


@pytest.mark.skipif(not os.path.isdir('/root/reference/trainer'), reason='needs the reference checkout (build container only)')
def test_reference_builder_builds_the_engine_from_its_own_config():
    """The seam end to end on the reference's side: its unmodified trainer/builder.py (which also imports trainer.trainer ->
    eval.counter, trainer.tester -> eval.coco_eval at import time) and its own north-star config dict, with the drop-in
    packages first on sys.path, construct this repo's model / post-process / COCOMetrics."""
    import subprocess
    from oracle import build_ref
    build_ref.write_stubs()
    code = '''
import sys
sys.path[:0] = [%r, %r, '/root/reference']
import torch
import config as config_module
from trainer import builder
import eval, model
cfg = getattr(config_module, 'orienmask_yolo_coco_544_anchor4_fpn_plus_infer')
cfg['model']['pretrained'] = None                                   # infer.py:79
m = builder.build(cfg['model'], builder.model_module)
p = builder.build_postprocess(cfg['postprocess'], device=torch.device('cuda:0'))
assert builder.tester_module.COCOMetrics.__module__ == 'orienmask_b200.coco_format'
assert eval.EvalCounter.__module__ == 'eval.counter'                # not replaced: the reference's own file
print(type(m).__module__, type(p).__module__, len(m.state_dict()), p.nms_thresh, p.conf_thresh, p.nms_pre, p.nms_post, p.grid_size)
''' % (os.path.join(ROOT, 'orienmask_b200', 'dropin'), build_ref.STUBS)
    out = subprocess.check_output([sys.executable, '-c', code], cwd='/tmp').decode().split(None, 7)
    assert out[:7] == ['orienmask_b200.model', 'orienmask_b200.postprocess', '524', '0.5', '0.005', '400', '100']
    assert out[7].strip() == '[(17, 17), (34, 34), (68, 68)]'


def test_launcher_puts_the_dropin_before_the_script_directory(tmp_path):
    """`python script.py` puts the script's directory first on sys.path, ahead of PYTHONPATH: a reference-like tree imports its own
    `model` package then.  `python -m orienmask_b200.dropin script.py` resolves the same import to the drop-in."""
    import subprocess
    ref = tmp_path / 'ref'
    (ref / 'model').mkdir(parents=True)
    (ref / 'model' / '__init__.py').write_text("OrienMaskYOLOFPNPlus = 'the reference class'\n")
    (ref / 'main.py').write_text("import sys, model\nprint(getattr(model.OrienMaskYOLOFPNPlus, '__module__', model.OrienMaskYOLOFPNPlus), sys.argv[1:])\n")
    dropin = os.path.join(ROOT, 'orienmask_b200', 'dropin')
    plain = subprocess.check_output([sys.executable, 'main.py', '-x'], cwd=str(ref), env=dict(os.environ, PYTHONPATH=dropin)).decode()
    assert 'the reference class' in plain                                          # PYTHONPATH alone is not enough
    via = subprocess.check_output([sys.executable, '-m', 'orienmask_b200.dropin', 'main.py', '-x', '1'], cwd=str(ref),
                                  env=dict(os.environ, PYTHONPATH=ROOT)).decode()
    assert via.split()[0] == 'orienmask_b200.model' and "['-x', '1']" in via
    rc = subprocess.run([sys.executable, '-m', 'orienmask_b200.dropin'], cwd=str(ref), env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True)
    assert rc.returncode == 2 and b'usage' in rc.stderr


@pytest.mark.skipif(not os.path.isfile('/root/reference/infer.py'), reason='needs the reference checkout (build container only)')
def test_reference_infer_py_runs_unchanged_up_to_the_first_forward(tmp_path):
    """The reference's own infer.py, unmodified, through the launcher: its config (as JSON with n_gpu = 0, so that everything up to
    the forward runs on this GPU-less box), its builder, this repo's model constructed and strictly loaded from a checkpoint with
    the reference's 524 keys, the reference's transform / pad / visualiser, this repo's post-process -- until infer.py:155 calls the
    model, which refuses the CPU tensor (the engine has no CPU path).  On the B200 box the same command runs through."""
    import json
    import subprocess
    from oracle import build_ref
    from orienmask_b200.synthetic import synthetic_state_dict
    build_ref.write_stubs()
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, build_ref.STUBS]))
    code = ("import json, sys; sys.path.insert(0, '/root/reference'); import config as C; "
            "cfg = dict(C.orienmask_yolo_coco_544_anchor4_fpn_plus_infer); cfg['n_gpu'] = 0; json.dump(cfg, open(sys.argv[1], 'w'))")
    cfg_file, weights = str(tmp_path / 'infer_cpu.json'), str(tmp_path / 'weights.pth')
    subprocess.check_call([sys.executable, '-c', code, cfg_file], cwd='/tmp', env=env)
    assert json.load(open(cfg_file))['model']['type'] == 'OrienMaskYOLOFPNPlus'
    torch.save({'state_dict': synthetic_state_dict(0)}, weights)                     # infer.py:82 accepts {'state_dict': ...}
    out = subprocess.run([sys.executable, '-m', 'orienmask_b200.dropin', 'infer.py', '-c', cfg_file, '-w', weights,
                          '-i', 'assets/000000163126.jpg'], cwd='/root/reference', env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert 'infer.py", line 155' in out.stderr and 'predictions = model(image)' in out.stderr, out.stderr[-2000:]
    assert 'orienmask_b200 runs on CUDA (sm_100a) only' in out.stderr and 'no CPU fallback' in out.stderr


@pytest.mark.skipif(not os.path.isfile('/root/reference/test.py'), reason='needs the reference checkout (build container only)')
def test_reference_test_py_runs_unchanged_up_to_the_first_forward(tmp_path):
    """The reference's own test.py, unmodified, through the launcher: build_tester reads the model config from the checkpoint,
    builds this repo's model and post-process by name, loads the 524 keys strictly, builds the reference's dataset / transform /
    dataloader on a one-image COCO-style set, and the reference's Tester (holding this repo's COCOMetrics) reaches
    trainer/tester.py:40 `predict = self.model(image)`, where the engine refuses the CPU tensor (n_gpu = 0 on this GPU-less box)."""
    import json
    import subprocess
    from oracle import build_ref
    from orienmask_b200.synthetic import synthetic_state_dict
    build_ref.write_stubs()
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, build_ref.STUBS]))
    (tmp_path / 'list.txt').write_text('000000163126.jpg\n')
    (tmp_path / 'anno.json').write_text(json.dumps({'000000163126.jpg': {'image_id': 163126, 'anno': {'bbox': [], 'cls': [], 'mask': []}}}))
    code = '''
import copy, json, sys, torch
sys.path.insert(0, '/root/reference')
import config as C
out = sys.argv[1]
cfg = copy.deepcopy(C.orienmask_yolo_coco_544_anchor4_fpn_plus_test)
cfg['n_gpu'], cfg['gt_file'] = 0, None
cfg['test_loader'].update(batch_size=1, num_workers=0)
cfg['test_loader']['dataset'].update(list_file=out + '/list.txt', image_dir='/root/reference/assets', anno_file=out + '/anno.json')
json.dump(cfg, open(out + '/test_cpu.json', 'w'))
json.dump(copy.deepcopy(C.orienmask_yolo_coco_544_anchor4_fpn_plus['model']), open(out + '/model.json', 'w'))
'''
    subprocess.check_call([sys.executable, '-c', code, str(tmp_path)], cwd='/tmp', env=env)
    model_cfg = json.load(open(str(tmp_path / 'model.json')))
    assert model_cfg['type'] == 'OrienMaskYOLOFPNPlus' and model_cfg['pretrained']            # build_tester must override it with None
    torch.save({'state_dict': synthetic_state_dict(0), 'config': {'model': model_cfg}}, str(tmp_path / 'ckpt.pth'))
    out = subprocess.run([sys.executable, '-m', 'orienmask_b200.dropin', 'test.py', '-c', str(tmp_path / 'test_cpu.json'),
                          '-w', str(tmp_path / 'ckpt.pth')], cwd='/root/reference', env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert 'trainer/tester.py", line 40' in out.stderr and 'predict = self.model(image)' in out.stderr, out.stderr[-2000:]
    assert 'orienmask_b200 runs on CUDA (sm_100a) only' in out.stderr


def test_shard_bounds_cover_batch():
    from orienmask_b200.sharding import shard_bounds
    for total in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from orienmask_b200.sharding import gather_detections, shard_bounds
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    det = torch.rand(4, 10, 5, generator=g)
    cls = torch.randint(0, 80, (4, 10), generator=g)
    cnt = torch.randint(0, 11, (4,), generator=g, dtype=torch.int32)
    lo, hi = shard_bounds(4, rank, world)
    d, c, n = gather_detections(det[lo:hi], cls[lo:hi], cnt[lo:hi])
    ok = torch.equal(d, det) and torch.equal(c, cls) and torch.equal(n, cnt)
    # ragged split: 5 images over 2 ranks (3 + 2); the collective stays fixed-size, the pad row is dropped on arrival
    det5, cls5, cnt5 = torch.rand(5, 10, 5, generator=g), torch.randint(0, 80, (5, 10), generator=g), torch.arange(5, dtype=torch.int32)
    lo, hi = shard_bounds(5, rank, world)
    d, c, n = gather_detections(det5[lo:hi], cls5[lo:hi], cnt5[lo:hi], total=5)
    ok = ok and torch.equal(d, det5) and torch.equal(c, cls5) and torch.equal(n, cnt5)
    try:
        gather_detections(det5[:1], cls5[:1], cnt5[:1], total=5)
        ok = False
    except ValueError:
        pass
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_detections_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(results) == [(0, True), (1, True)]


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` prints one JSON line with the keys the driver reads (CPU only: the oracle port)."""
    import json
    import subprocess
    import sys
    from tests.common import ROOT
    out = subprocess.run([sys.executable, 'bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'], cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['cpu_baseline']['cores'] >= 1
