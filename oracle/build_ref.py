"""TEST INFRASTRUCTURE ONLY -- recipe that makes the *unmodified* reference importable here.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu-baseline / ``--impl reference`` legs may import it.

What this does (outputs only into ``oracle/_ref/``, which is git-ignored):

* compiles the reference's native greedy NMS from the source file where it lies
  (``/root/reference/eval/src/nms_cpu.cpp``) into ``oracle/_ref/ref_nms_cpu*.so``.  The file does
  not build against torch 2.11 as-is (``dets.type()`` is no longer convertible to a ScalarType at
  ``eval/src/nms_cpu.cpp:67``), so the translation unit handed to the compiler is a one-token
  patched temporary (``dets.type()`` -> ``dets.scalar_type()`` on that line) written to
  ``oracle/_ref/build``; no reference source is ever committed.
* writes empty stub modules for the packages the reference imports but this image lacks
  (``torchsummary``, ``prettytable``, ``tensorboardX``, ``pycocotools``, and ``matplotlib`` for ``infer.py`` itself) to
  ``oracle/_ref/stubs``.

``import_reference()`` then imports the reference's own ``config`` / ``model`` / ``eval`` packages
from ``/root/reference`` (only possible in the build container -- the GPU box has no reference).
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
STUBS = os.path.join(OUT, 'stubs')
# A byte-for-byte copy of the reference checkout that travels to the GPU box with the repo snapshot (git-ignored, NOT
# gpurun-ignored; made by ship_reference() in the build container, where /root/reference exists).  Only `-m gpu` TESTS that run
# the reference's own infer.py / test.py / nn.Module on the B200 read it; the product, smoke() and bench.py never do.
SHIPPED = os.path.join(os.path.dirname(HERE), 'baseline', '_ref', 'reference')


def _default_ref_root():
    env = os.environ.get('ORIENMASK_REFERENCE')
    if env:
        return env
    if os.path.isfile('/root/reference/eval/src/nms_cpu.cpp'):
        return '/root/reference'
    return SHIPPED


REF_ROOT = _default_ref_root()

_STUB_SOURCES = {
    'torchsummary.py': 'def summary(*a, **k):\n    pass\n',
    'prettytable.py': 'class PrettyTable:\n    pass\n',
    'tensorboardX.py': 'class SummaryWriter:\n    pass\n',
    'pycocotools/__init__.py': '',
    'pycocotools/mask.py': '',
    # Stand-ins with pycocotools' call surface, enough for the reference's test.py / Tester to run to its last line on a box
    # without the real package: ground truth and results are loaded, NO AP is computed -- every statistic is -1, which is what
    # pycocotools itself reports when nothing can be matched.  Test infrastructure only.
    'pycocotools/coco.py': """import json


class COCO:
    def __init__(self, annotation_file=None):
        self.dataset = json.load(open(annotation_file)) if annotation_file else {'images': [], 'annotations': [], 'categories': []}
        self.anns = list(self.dataset.get('annotations', []))

    def getImgIds(self):
        return [im['id'] for im in self.dataset.get('images', [])]

    def getCatIds(self):
        return [c['id'] for c in self.dataset.get('categories', [])]

    def loadRes(self, res_file):
        res = COCO()
        res.dataset = dict(self.dataset)
        res.anns = json.load(open(res_file)) if isinstance(res_file, str) else list(res_file)
        return res
""",
    'pycocotools/cocoeval.py': """import numpy as np


class COCOeval:
    def __init__(self, coco_gt=None, coco_dt=None, iouType='segm'):
        self.gt, self.dt, self.iouType = coco_gt, coco_dt, iouType
        self.stats, self.eval = None, None

    def evaluate(self):
        pass

    def accumulate(self):
        k = max(1, len(self.gt.getCatIds()))
        self.eval = {'precision': -np.ones((10, 101, k, 4, 3)), 'recall': -np.ones((10, k, 4, 3))}

    def summarize(self):
        self.stats = -np.ones(12)
        print('pycocotools stand-in: %d %s results loaded, no AP computed' % (len(self.dt.anns), self.iouType))
""",
    'matplotlib/__init__.py': '',                 # only the reference's infer.py itself imports it (infer.py:9)
    'matplotlib/pyplot.py': '',
}


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'eval', 'src', 'nms_cpu.cpp'))


def write_stubs():
    for rel, src in _STUB_SOURCES.items():
        path = os.path.join(STUBS, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, 'w') as f:
            f.write(src)


def ship_reference():
    """Copy the reference checkout into baseline/_ref/reference (git-ignored; it ships to the GPU box with the snapshot) so that
    the `-m gpu` drop-in tests can run the reference's own, unmodified infer.py / test.py / nn.Module there.  No-op when
    /root/reference is absent (i.e. on the GPU box itself).  Returns the path of the copy or None."""
    src = '/root/reference'
    if not os.path.isfile(os.path.join(src, 'infer.py')):
        return SHIPPED if os.path.isfile(os.path.join(SHIPPED, 'infer.py')) else None
    import shutil
    if os.path.isdir(SHIPPED):
        shutil.rmtree(SHIPPED)
    shutil.copytree(src, SHIPPED, ignore=shutil.ignore_patterns('.git', '__pycache__', '*.pyc', '*.log', 'framework.png'))
    return SHIPPED


def reference_root():
    """The reference tree usable in this process (/root/reference in the build container, the shipped copy on the GPU box) or None."""
    for cand in (os.environ.get('ORIENMASK_REFERENCE'), '/root/reference', SHIPPED):
        if cand and os.path.isfile(os.path.join(cand, 'infer.py')):
            return cand
    return None


def _find_built():
    if not os.path.isdir(OUT):
        return None
    for name in os.listdir(OUT):
        if name.startswith('ref_nms_cpu') and name.endswith('.so'):
            return os.path.join(OUT, name)
    return None


def build_ref_nms(verbose=False):
    """Compile the reference nms_cpu.cpp (one-token patched temp copy) -> oracle/_ref/ref_nms_cpu.so."""
    built = _find_built()
    if built:
        return built
    if not reference_available():
        return None
    from torch.utils.cpp_extension import load
    bdir = os.path.join(OUT, 'build')
    os.makedirs(bdir, exist_ok=True)
    with open(os.path.join(REF_ROOT, 'eval', 'src', 'nms_cpu.cpp')) as f:
        text = f.read()
    patched = text.replace('AT_DISPATCH_FLOATING_TYPES(dets.type(),', 'AT_DISPATCH_FLOATING_TYPES(dets.scalar_type(),')
    assert patched != text, 'reference nms_cpu.cpp changed: patch point not found'
    tmp = os.path.join(bdir, 'ref_nms_cpu_tu.cpp')
    with open(tmp, 'w') as f:
        f.write(patched)
    load(name='ref_nms_cpu', sources=[tmp], build_directory=bdir, verbose=verbose)
    import shutil
    for name in os.listdir(bdir):
        if name.endswith('.so'):
            shutil.copy(os.path.join(bdir, name), os.path.join(OUT, 'ref_nms_cpu.so'))
    return _find_built()


def load_ref_nms():
    """Return the compiled reference module exposing ``nms(dets[N,5], thr) -> LongTensor`` or None."""
    path = build_ref_nms()
    if path is None:
        return None
    import torch  # noqa: F401  (must be loaded before the extension)
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_nms_cpu', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    """Import the reference's own packages (build container only). Returns (config, model, builder)."""
    if not reference_available():
        raise RuntimeError('reference tree not present at %s' % REF_ROOT)
    write_stubs()
    nms_cpu = load_ref_nms()
    for m in list(sys.modules):
        if m.split('.')[0] in ('eval', 'model', 'config', 'trainer', 'utils', 'data', 'optim'):
            del sys.modules[m]
    sys.path.insert(0, STUBS)
    sys.path.insert(0, REF_ROOT)
    sys.modules['eval.nms_cpu'] = nms_cpu
    sys.modules['eval.nms_cuda'] = types.ModuleType('eval.nms_cuda')
    config = importlib.import_module('config')
    model = importlib.import_module('model')
    builder = importlib.import_module('trainer.builder')
    return config, model, builder


def release_reference():
    """Undo the sys.path / sys.modules changes of import_reference()."""
    for p in (STUBS, REF_ROOT):
        while p in sys.path:
            sys.path.remove(p)
    for m in list(sys.modules):
        if m.split('.')[0] in ('eval', 'model', 'config', 'trainer', 'utils', 'data', 'optim'):
            del sys.modules[m]


if __name__ == '__main__':
    write_stubs()
    print('ref nms:', build_ref_nms(verbose=True))
