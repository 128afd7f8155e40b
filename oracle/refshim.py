"""Imports the UNMODIFIED reference for golden generation, the host-logic tests, the `-m gpu` tests that
drive the reference's own planners through the CUDA drop-ins, and bench.py's CPU arm.
TEST / BASELINE INFRASTRUCTURE ONLY -- the product never imports this module.

Where it comes from: the read-only checkout (/root/reference) in the build container; on the GPU box, where
that path does not exist, the verbatim copy that `make -C oracle ref` staged under oracle/_ref/ (git-ignored,
shipped with the tree like the built libraries).

pomp.example_problems does `from OpenGL.GL import *`; PyOpenGL is not installed, so no-op stub
modules are injected (drawing code is never called).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    env = os.environ.get("POMP_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference", os.path.join(_HERE, "_ref")]:
        if cand and os.path.isdir(os.path.join(cand, "pomp")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pomp"))


class _Any(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return lambda *a, **kw: None


def load():
    """Returns the `pomp` package of the reference checkout."""
    if not available():
        raise ImportError("reference checkout not found at %s" % REFERENCE_ROOT)
    if "OpenGL" not in sys.modules:
        og = _Any("OpenGL")
        sys.modules["OpenGL"] = og
        for sub in ("GL", "GLU", "GLUT"):
            m = _Any("OpenGL." + sub)
            m.__all__ = []
            sys.modules["OpenGL." + sub] = m
            setattr(og, sub, m)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pomp  # noqa: F401
    import pomp.structures.nearestneighbors  # noqa: F401
    import pomp.spaces.metric  # noqa: F401
    import pomp.spaces.edgechecker  # noqa: F401
    import pomp.planners.allplanners  # noqa: F401
    import pomp.example_problems.geometric  # noqa: F401
    import pomp.example_problems.dubins  # noqa: F401
    import pomp.example_problems.pendulum  # noqa: F401
    import pomp.example_problems.doubleintegrator  # noqa: F401
    import pomp.example_problems.flappy  # noqa: F401
    return pomp
