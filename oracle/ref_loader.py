"""TEST INFRASTRUCTURE ONLY -- loader for the unmodified reference.

Imports ``/root/reference`` (bezier.py, optimization.py, gjk/gjk.py) in-process
so that golden vectors can be generated from the reference itself and so the
numpy/C restatements under ``oracle/`` can be pinned against it.  Nothing in
the product package may import this module; ``/root/reference`` only exists in
the authoring container (never on the GPU box), so every caller must treat
``available() == False`` as "skip".

What is needed to import the reference:
  * NUMBA_CACHE_DIR / PYTHONDONTWRITEBYTECODE so nothing is written under the
    read-only reference tree (its njit functions use ``cache=True``);
  * stub ``matplotlib`` modules (bezier.py:15-16, optimization.py:9 import it;
    matplotlib is not installed here);
  * two shims for the minDist/collCheck family, which is broken at HEAD:
      - ``bezier.gjkNew`` is never imported (bezier.py:21-22 are commented out)
      - ``BezierParams.__init__`` needs ``.ndim`` but ``_minDist`` passes a
        list (bezier.py:1304, bezier.py:58).
No reference file is modified or copied.
"""
import importlib
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("BEZ_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bezier.py"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    pyplot = types.ModuleType("matplotlib.pyplot")
    anim = types.ModuleType("matplotlib.animation")
    mplt = types.ModuleType("mpl_toolkits")
    m3d = types.ModuleType("mpl_toolkits.mplot3d")
    m3d.Axes3D = object
    mpl.pyplot = pyplot
    mpl.animation = anim
    mplt.mplot3d = m3d
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = pyplot
    sys.modules["matplotlib.animation"] = anim
    sys.modules["mpl_toolkits"] = mplt
    sys.modules["mpl_toolkits.mplot3d"] = m3d


def load():
    """Returns a namespace with the reference modules: .bezier .optimization .gjk"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    os.environ.setdefault("NUMBA_CACHE_DIR",
                          os.path.join(tempfile.gettempdir(), "bez_numba_cache"))
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    _stub_matplotlib()

    # The product package deliberately re-uses the reference's module names for
    # its drop-in mirror (as sub-modules of the package), so there is no clash
    # with the top-level names imported here.
    saved = list(sys.path)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import numpy as np
        ref_bezier = importlib.import_module("bezier")
        ref_opt = importlib.import_module("optimization")
        ref_gjk = importlib.import_module("gjk.gjk")
    finally:
        sys.path[:] = saved

    # shim (a): gjkNew is never imported by bezier.py
    ref_bezier.gjkNew = ref_gjk.gjkNew

    # shim (b): list-of-rows constructor used by _minDist & friends
    if not getattr(ref_bezier.BezierParams, "_oracle_shimmed", False):
        orig_init = ref_bezier.BezierParams.__init__

        def patched(self, cpts=None, tau=None, t0=0.0, tf=1.0):
            if cpts is not None and not isinstance(cpts, np.ndarray):
                cpts = np.array(cpts, dtype=float)
            orig_init(self, cpts=cpts, tau=tau, t0=t0, tf=tf)

        ref_bezier.BezierParams.__init__ = patched
        ref_bezier.BezierParams._oracle_shimmed = True

    ns = types.SimpleNamespace(bezier=ref_bezier, optimization=ref_opt, gjk=ref_gjk)
    _loaded = ns
    return ns


def load_example(name):
    """Imports Examples/<name>.py as a module (drivers are under __main__)."""
    load()
    path = os.path.join(REFERENCE_ROOT, "Examples", name + ".py")
    saved = list(sys.path)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        spec = importlib.util.spec_from_file_location("ref_example_" + name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved
    return mod
