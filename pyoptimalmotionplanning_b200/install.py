"""Wiring the drop-ins into an unmodified pomp checkout.

pomp selects its hot-path objects through three existing knobs:
  * `nearestNeighborMethod=` -> NearestNeighbors(metric, method)   (kinodynamicplanner.py:298-299,
     rrtstarplanner.py:37-38,232-234); the planner modules bind the class by name through
     `from ..structures.nearestneighbors import *`,
  * `checker=` / `metric=` of makePlanner                          (allplanners.py:29-35,66-68),
  * planner.setControlSelector(selector)                           (kinodynamicplanner.py:341).

install(pomp) rebinds NearestNeighbors in the three modules that hold the name and leaves
everything else untouched; uninstall(pomp) restores the originals.
"""
import importlib

from . import structures, spaces

_MODULES = ("pomp.structures.nearestneighbors", "pomp.planners.kinodynamicplanner", "pomp.planners.rrtstarplanner")
_saved = {}


def install(nn_factory=None):
    """Route every NearestNeighbors(...) the planners construct to the GPU drop-in.
    nn_factory(metric, method) may be given to customise construction (tests use it)."""
    def make(metric, method="b200"):
        if nn_factory is not None:
            return nn_factory(metric, method)
        return structures.NearestNeighbors(metric, method)

    for name in _MODULES:
        mod = importlib.import_module(name)
        if name not in _saved:
            _saved[name] = mod.NearestNeighbors
        mod.NearestNeighbors = make


def uninstall():
    for name, cls in _saved.items():
        importlib.import_module(name).NearestNeighbors = cls
    _saved.clear()


def make_checker(space, resolution=0.01, **kw):
    """checker= argument for pomp.planners.allplanners.makePlanner."""
    return spaces.EpsilonEdgeChecker(space, resolution, **kw)
