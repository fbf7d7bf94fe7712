"""Shared by the drop-in ``model`` / ``eval`` packages: chain a shadowing package to the package it shadows.

With ``orienmask_b200/dropin`` first on ``sys.path`` the reference's ``infer.py`` / ``test.py`` import *these* ``model`` and
``eval`` packages.  They replace the hot-path names only; the reference's own modules import further submodules of the
same packages that have nothing to do with the path (``trainer/trainer.py:9`` needs ``eval.counter`` at import time,
the training loss needs ``eval.base``).  Appending the shadowed package's directory to ``__path__`` lets every submodule
that is not replaced here resolve to the reference's own file, unchanged.
"""
import os
import sys


def chain_to_shadowed(package_name, package_path, own_dir):
    own = os.path.realpath(own_dir)
    for entry in list(sys.path):
        cand = os.path.join(entry or os.getcwd(), package_name)
        if os.path.realpath(cand) != own and os.path.isfile(os.path.join(cand, '__init__.py')) and cand not in package_path:
            package_path.append(cand)
            return cand
    return None
