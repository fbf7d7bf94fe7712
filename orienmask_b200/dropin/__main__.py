"""Run one of the reference's own entry points (``infer.py``, ``test.py``) unchanged on the B200 engine:

    cd /path/to/OrienMask
    PYTHONPATH=/path/to/this/repo python -m orienmask_b200.dropin infer.py -c orienmask_yolo_coco_544_anchor4_fpn_plus_infer -w ... -i ...

Why a launcher: ``python infer.py`` puts the script's directory -- the reference root, which holds the real ``model`` and
``eval`` packages -- at ``sys.path[0]``, *before* anything named in ``PYTHONPATH``, so a drop-in directory on ``PYTHONPATH``
alone is never reached.  This module puts the drop-in packages first and the reference root second, then executes the script as
``__main__`` with its own argument vector; not one line of the reference is modified.  Packages the drop-in does not replace
(``config``, ``trainer``, ``data``, ``utils``, and the submodules ``eval.counter`` / ``eval.base`` ..., see ``_chain.py``) are
the reference's.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ('-h', '--help'):
        sys.stderr.write('usage: python -m orienmask_b200.dropin <reference script.py> [its arguments...]\n')
        return 2
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        sys.stderr.write('orienmask_b200.dropin: no such script: %s\n' % argv[0])
        return 2
    here = os.path.dirname(os.path.abspath(__file__))
    ref_root = os.path.dirname(script)
    for p in (ref_root, here):                       # final order: drop-in, reference root, everything else
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name='__main__')
    return 0


if __name__ == '__main__':
    sys.exit(main())
