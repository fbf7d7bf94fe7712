"""Golden fixture for the pre-process, made by the UNMODIFIED reference in this container.

    python tests/golden/make_golden_prep.py          (build container only: needs /root/reference)

Runs the reference's own ``data.transform.FastCOCOTransform`` (use_cuda=False) with the infer pipeline shape
(``config/base.py:158-164``: Resize + Normalize(mean 0, std 255)) and a ShortEdgeResize variant, then the
``pad`` function of ``infer.py:21-32`` (extracted from the file by ``ast`` because importing infer.py needs
matplotlib), on seeded uint8 images, and stores inputs and outputs in ``prep_small.npz``.
"""
import ast
import math  # noqa: F401  (used by the extracted function)
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import build_ref  # noqa: E402


def reference_pad():
    src = open(os.path.join(build_ref.REF_ROOT, 'infer.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'pad'][0]
    ns = {'math': math, 'F': torch.nn.functional}
    exec(compile(ast.Module([fn], []), 'infer.py', 'exec'), ns)
    return ns['pad']


def main():
    build_ref.import_reference()
    import data.transform as T
    pad = reference_pad()
    rng = np.random.default_rng(5)
    d = {}
    cases = {
        'resize': (rng.integers(0, 256, (2, 37, 50, 3), dtype=np.uint8),
                   [T.FastCOCOTransform.Resize(size=(64, 96)), T.FastCOCOTransform.Normalize(mean=(0, 0, 0), std=(255, 255, 255))]),
        'short': (rng.integers(0, 256, (1, 37, 50, 3), dtype=np.uint8),
                  [T.FastCOCOTransform.ShortEdgeResize(short_length=48, max_size=80),
                   T.FastCOCOTransform.Normalize(mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375))]),
        'down': (rng.integers(0, 256, (1, 150, 211, 3), dtype=np.uint8),
                 [T.FastCOCOTransform.Resize(size=(64, 64)), T.FastCOCOTransform.Normalize(mean=(0, 0, 0), std=(255, 255, 255))]),
    }
    for name, (img, pipeline) in cases.items():
        tr = T.FastCOCOTransform(pipeline, use_cuda=False)
        x = tr(torch.tensor(img, dtype=torch.float32))             # infer.py:148-149
        y, info = pad(x)                                            # infer.py:150
        print(name, tuple(img.shape), '->', tuple(x.shape), '->', tuple(y.shape), info)
        d[name + '_in'] = img
        d[name + '_out'] = y.numpy()
        d[name + '_pad'] = np.asarray(info, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, 'prep_small.npz'), **d)
    print(os.path.getsize(os.path.join(HERE, 'prep_small.npz')))


if __name__ == '__main__':
    main()
