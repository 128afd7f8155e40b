"""Imports the UNMODIFIED reference (read-only checkout) for golden generation and CPU-only
host-logic tests.  TEST INFRASTRUCTURE ONLY; the reference does not exist on the GPU box.

pomp.example_problems does `from OpenGL.GL import *`; PyOpenGL is not installed, so no-op stub
modules are injected (drawing code is never called).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("POMP_REFERENCE_ROOT", "/root/reference")


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
