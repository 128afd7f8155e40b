"""pomp-b200: the data-parallel hot path of pyOptimalMotionPlanning on B200 (sm_100a).

Drop-ins (same names and call signatures as the reference):
    NearestNeighbors, EpsilonEdgeChecker, BatchedRandomControlSelector, DeviceControlSpace,
    DeviceSpace
plus explicit descriptors (MetricSpec, SpaceSpec, DynamicsSpec) and a small kinematic RRT driver
(rrt.RRT / rrt.RepeatedRRT) that reproduces the reference planner's trees under a fixed seed with
one fused launch per iteration.

Compute happens only in csrc/libpomp_b200.so (hand-written CUDA behind a C ABI, see
include/pomp_b200.h).  There is no CPU fallback.
"""
from .descriptors import (MetricSpec, SpaceSpec, DynamicsSpec, metric_from_pomp, space_from_pomp,
                          dynamics_from_pomp)
from .structures import NearestNeighbors, ProjectionHashDensity
from .spaces import DeviceSpace, EpsilonEdgeChecker, DeviceControlSpace, BatchedRandomControlSelector
from .install import install, uninstall, make_checker
from .sst import SSTStarForest
from .roadmap import DeviceRoadmap
from ._abi import PompError, load_library

__all__ = ["MetricSpec", "SpaceSpec", "DynamicsSpec", "metric_from_pomp", "space_from_pomp", "dynamics_from_pomp",
           "NearestNeighbors", "ProjectionHashDensity", "DeviceSpace", "EpsilonEdgeChecker", "DeviceControlSpace",
           "BatchedRandomControlSelector", "SSTStarForest", "DeviceRoadmap", "install", "uninstall", "make_checker", "PompError", "load_library"]
