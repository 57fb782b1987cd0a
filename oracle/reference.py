"""Loaders for the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

* ``subg_acc()``: the reference's C extension, compiled from
  /root/reference/subg_acc/subg_acc.c into oracle/_ref/ by ``make -C oracle ref``
  (the built file travels to the GPU box; the sources never enter this repo).
* ``train()`` / ``pprgo()``: the reference's Python modules imported from
  /root/reference.  Only possible where that tree exists (this container), so
  they are used solely by tests/golden/make_golden.py and by the
  ``reference``-marked CPU tests; nothing that runs on the GPU box needs them.
"""
from __future__ import annotations

import glob
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SUBG_REFERENCE_ROOT", "/root/reference")


def have_reference_tree() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "train.py"))


def subg_acc():
    """Import the compiled reference extension; None if it has not been built."""
    hits = glob.glob(os.path.join(_HERE, "_ref", "subg_acc*.so"))
    if not hits:
        return None
    name = "_reference_subg_acc"
    if name in sys.modules:
        return sys.modules[name]
    # the extension's init symbol is PyInit_subg_acc, so the spec name must be subg_acc
    spec = importlib.util.spec_from_file_location("subg_acc", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


def _import_from_reference(modname: str, relpath: str):
    if not have_reference_tree():
        return None
    key = f"_reference_{modname}"
    if key in sys.modules:
        return sys.modules[key]
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # sampler/pprgo.py:78 uses the alias removed in numpy 1.24
    spec = importlib.util.spec_from_file_location(key, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[key] = mod
    return mod


def train():
    """/root/reference/train.py (gather, bgather, pgather, hgather)."""
    return _import_from_reference("train", "train.py")


def pprgo():
    """/root/reference/sampler/pprgo.py (topk_ppr_matrix, calc_ppr_topk_parallel).
    Imported under its own package name ``sampler.pprgo`` because its numba
    on-disk cache (``cache=True``) pickles that module path."""
    if not have_reference_tree():
        return None
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # sampler/pprgo.py:78 uses the alias removed in numpy 1.24
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        return importlib.import_module("sampler.pprgo")
    finally:
        sys.path.remove(REFERENCE_ROOT)
