"""Device-resident k-nearest-neighbour roadmap (BASELINE config 3: `Kink 2D roadmap: 4M samples, batched edge
visibility checks vs 10K box obstacles`).

The reference exports what its planners built as a graph (V, E) -- TreePlanner.getRoadmap,
pomp/planners/kinodynamicplanner.py:211-238: V a list of states, E triples (i, j, u).  A roadmap over a fixed sample
set is the same three calls of the hot path per sample that the planners make one at a time --
`space.feasible(x)` (example_problems/geometric.py:56-60), `nn.knearest(x, k)` (structures/nearestneighbors.py:98-116)
and `edgeChecker.feasible(space.interpolator(x, y))` (spaces/edgechecker.py:20-30) -- so here they run as ONE device
pipeline over all samples (csrc/roadmap.cu, `pomp_roadmap_build`): nothing returns to the host but the edge count
(and, on request, the graph itself).
"""
import ctypes as C

import numpy as np

from . import _abi
from .backend import check, _ptr


class DeviceRoadmap(object):
    """nn: backend.CudaNN holding the samples; space: backend.CudaSpace.  Both on the same GPU."""

    def __init__(self, nn, space):
        self.nn, self.space = nn, space
        self.lib = _abi.load_library()
        self.node_ok = None
        self.edges = None
        self.n_candidates = 0

    def build(self, k, resolution=0.01, undirected=True, capacity=None):
        """Builds the roadmap and copies (node_ok, edges) to the host.  Returns the number of feasible edges."""
        n = int(self.nn.size())
        cap = int(capacity) if capacity is not None else n * k
        node_ok = np.empty(max(n, 1), np.uint8)
        edges = np.empty((max(cap, 1), 2), np.int32)
        ne, nc = C.c_int64(0), C.c_int64(0)
        check(self.lib, self.lib.pomp_roadmap_build_h(self.nn._h, self.space._h, k, float(resolution),
                                                       1 if undirected else 0, _ptr(node_ok), _ptr(edges), cap,
                                                       C.byref(ne), C.byref(nc)))
        if ne.value > cap:
            return self.build(k, resolution, undirected, capacity=ne.value)
        self.node_ok = node_ok[:n].astype(bool)
        self.edges = edges[:ne.value].copy()
        self.n_candidates = int(nc.value)
        return int(ne.value)

    def build_device(self, k, resolution, undirected, node_ok_ptr, edges_ptr, capacity, stream=None):
        """Device-pointer form (outputs stay in HBM): returns (n_edges, n_candidates)."""
        ne, nc = C.c_int64(0), C.c_int64(0)
        check(self.lib, self.lib.pomp_roadmap_build(self.nn._h, self.space._h, k, float(resolution),
                                                     1 if undirected else 0, node_ok_ptr, edges_ptr, int(capacity),
                                                     C.byref(ne), C.byref(nc), stream))
        return int(ne.value), int(nc.value)

    def getRoadmap(self, points):
        """(V, E) in TreePlanner.getRoadmap's shape: V the states, E triples (i, j, None) (straight edges carry no
        control).  `points`: the samples as the caller holds them (the node set itself stays on the device)."""
        V = [list(map(float, p)) for p in points]
        E = [(int(i), int(j), None) for i, j in self.edges]
        return V, E

    def to_csv(self, points, vertices_file, edges_file):
        """Two plain CSV files (`x0,..,feasible` per vertex; `i,j` per edge)."""
        pts = np.asarray(points, dtype=np.float64)
        with open(vertices_file, "w") as f:
            for p, ok in zip(pts, self.node_ok):
                f.write(",".join(repr(float(v)) for v in p) + "," + ("1" if ok else "0") + "\n")
        with open(edges_file, "w") as f:
            for i, j in self.edges:
                f.write("%d,%d\n" % (i, j))
