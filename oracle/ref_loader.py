"""Import the UNMODIFIED reference modules from oracle/_ref/ (built by oracle/make_ref.sh).  TEST INFRASTRUCTURE ONLY.

The reference uses bare-module imports (``from krylow import ...``, gauss_newton_krylow.py:3-9), the same names that
``gauss_newton_via_generalized_krylov_subspaces_b200.install_flat_names()`` registers for the B200 mirror.  To keep
the two apart, the reference modules are imported with oracle/_ref first on ``sys.path`` and then taken OUT of
``sys.modules`` again; the returned namespace holds the only references to them.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULES = ("regression_result", "armijo_goldstein", "krylow", "gauss_newton", "gauss_newton_krylow",
           "bratu_pde_problem", "rosenbrock_problem", "benchmark")


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, m + ".py")) for m in MODULES)


_cached = None


def load_reference():
    """-> namespace with the reference's modules as attributes (``ref.gauss_newton_krylow.gauss_newton_krylow`` ...),
    or None when oracle/_ref has not been built."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    saved = {m: sys.modules.pop(m) for m in MODULES if m in sys.modules}
    sys.path.insert(0, REF_DIR)
    ns = types.SimpleNamespace()
    try:
        for m in MODULES:
            setattr(ns, m, importlib.import_module(m))
    finally:
        sys.path.remove(REF_DIR)
        for m in MODULES:
            sys.modules.pop(m, None)
        sys.modules.update(saved)
    ns.dir = REF_DIR
    _cached = ns
    return ns
