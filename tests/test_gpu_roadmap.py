"""SURVEY 8 row f4 on the GPU: the device-resident k-NN roadmap (csrc/roadmap.cu, BASELINE config 3) against the
oracle's feasible / knearest / edge_check composed on the host, and the results CSV of the lock-step trial harness
(rrt.testForest, the layout of pomp/planners/test.py:8-52) written from a run on the CUDA back end."""
import csv

import numpy as np
import pytest

from oracle import oracle
from pyoptimalmotionplanning_b200 import MetricSpec, SpaceSpec
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _field(rng, nbox):
    c = rng.random((nbox, 2))
    hs = 0.004 + rng.random((nbox, 2)) * 0.012
    kink = [[0.3, 0, 0.51, 0.19], [0.51, 0, 0.7, 0.29], [0.3, 0.21, 0.49, 0.7], [0.49, 0.31, 0.7, 0.7]]
    return SpaceSpec([0, 0], [1, 1], boxes=np.concatenate([np.array(kink), np.concatenate([c - hs, c + hs], 1)]))


def _host_roadmap(sp, pts, k, res, undirected):
    """The same roadmap from the oracle's per-call functions, composed the way a host loop over the samples would."""
    m = MetricSpec.euclidean(2)
    n = len(pts)
    node_ok = oracle.feasible(sp, pts).astype(bool)
    ki, kd, kc = oracle.nn_knearest(m, pts, pts, k + 1)
    lists = [set(r.tolist()) for r in ki]
    cand = []
    for i in range(n):
        row = ki[i].tolist()
        self_at = row.index(i) if i in row else k
        for t, j in enumerate(row):
            if t == self_at:
                continue
            if undirected and j < i and i in lists[j]:
                continue
            cand.append((i, j))
    cand = np.array(cand, np.int64).reshape(-1, 2)
    ok = oracle.edge_check(sp, pts[cand[:, 0]], pts[cand[:, 1]], res).astype(bool)
    return node_ok, cand[ok], len(cand)


@pytest.mark.parametrize("undirected", [True, False])
def test_device_roadmap_equals_host_composition(undirected):
    from pyoptimalmotionplanning_b200.backend import CudaNN, CudaSpace
    from pyoptimalmotionplanning_b200.roadmap import DeviceRoadmap
    rng = np.random.default_rng(12)
    sp = _field(rng, 600)
    pts = rng.random((30000, 2))
    pts[100] = pts[50]                      # an exact duplicate: a node that is not first in its own list
    nn = CudaNN(MetricSpec.euclidean(2), 0)
    nn.set(pts)
    rm = DeviceRoadmap(nn, CudaSpace(sp, 0))
    for k, res in ((6, 0.01), (3, 0.002)):
        ne = rm.build(k, res, undirected)
        node_ok, edges, ncand = _host_roadmap(sp, pts, k, res, undirected)
        assert (rm.node_ok == node_ok).all()
        assert ne == len(edges) and (rm.edges == edges).all()        # same edges in the same order
        if not undirected:
            assert rm.n_candidates == ncand == len(pts) * k
    # too small a capacity is reported, never overrun; the wrapper retries with the right size
    assert rm.build(6, 0.01, undirected, capacity=10) == len(_host_roadmap(sp, pts, 6, 0.01, undirected)[1])
    V, E = rm.getRoadmap(pts)
    assert len(V) == len(pts) and E[0][2] is None and len(E) == len(rm.edges)


def test_roadmap_csv_export(tmp_path):
    from pyoptimalmotionplanning_b200.backend import CudaNN, CudaSpace
    from pyoptimalmotionplanning_b200.roadmap import DeviceRoadmap
    rng = np.random.default_rng(3)
    sp = _field(rng, 100)
    pts = rng.random((20000, 2))
    nn = CudaNN(MetricSpec.euclidean(2), 0)
    nn.set(pts)
    rm = DeviceRoadmap(nn, CudaSpace(sp, 0))
    ne = rm.build(4, 0.01)
    vf, ef = str(tmp_path / "v.csv"), str(tmp_path / "e.csv")
    rm.to_csv(pts, vf, ef)
    vrows = list(csv.reader(open(vf)))
    erows = list(csv.reader(open(ef)))
    assert len(vrows) == len(pts) and len(erows) == ne
    assert [float(vrows[7][0]), float(vrows[7][1])] == pts[7].tolist()          # repr round-trips the doubles
    assert [int(v[2]) for v in vrows] == oracle.feasible(sp, pts).astype(int).tolist()


def test_forest_results_csv_on_cuda_has_the_reference_layout(tmp_path):
    """rrt.testForest on the CUDA back end: pomp/planners/test.py's CSV (header, a row per change of a trial's best
    cost, a closing row per trial); trial 0 under seed 0 reproduces the golden cost trace of the reference's r-rrt."""
    from pyoptimalmotionplanning_b200.rrt import RepeatedRRTForest, testForest
    g1 = G.load("planners.json")["G1_rrrt_bugtrap_bruteforce"]
    rf = RepeatedRRTForest(G.space_spec(g1["space"]), g1["start"], [0, 5, 9], g1["goal"], g1["goal_radius"],
                           max_step=0.15, resolution=0.01)
    fn = str(tmp_path / "rrrt.csv")
    rows = testForest(rf, 700, fn, trial_offset=4)
    lines = open(fn).read().splitlines()
    assert lines[0] == "trial,plan iters,plan time,best cost"
    parsed = list(csv.reader(lines[1:]))
    assert all(len(r) == 4 for r in parsed) and len(parsed) == len(rows)
    assert [int(r[0]) for r in parsed] == sorted(int(r[0]) for r in parsed) and {r[0] for r in parsed} == {"4", "5", "6"}
    t0 = [(int(r[1]), float(r[3])) for r in parsed if r[0] == "4"]
    want = [(c[0], c[1]) for c in g1["change_points"] if c[0] <= 700]
    assert t0[:-1] == want and t0[-1] == (700, want[-1][1])


def test_indexed_edges_with_node_flags_equal_plain_indexed_edges():
    """pomp_edge_check_indexed_nodes (feasibility once per node, the edges gather the flags) returns the booleans of
    pomp_edge_check_indexed and of the oracle -- short roadmap edges, long edges, infeasible and out-of-bounds nodes."""
    import torch
    from pyoptimalmotionplanning_b200.backend import CudaSpace
    rng = np.random.default_rng(21)
    sp = _field(rng, 800)
    space = CudaSpace(sp, 0)
    nodes = rng.random((50000, 2)) * 1.04 - 0.02                   # a few nodes outside the bounds
    ij = np.concatenate([np.stack([np.arange(50000), (np.arange(50000) + 1) % 50000], 1),        # long edges
                         np.stack([rng.integers(0, 50000, 100000), rng.integers(0, 50000, 100000)], 1)]).astype(np.int32)
    near = np.argsort(nodes[:2000, 0])                              # short edges between x-neighbours
    ij = np.concatenate([ij, np.stack([near[:-1], near[1:]], 1).astype(np.int32)])
    d_nodes, d_ij = torch.from_numpy(nodes).cuda(), torch.from_numpy(ij).cuda()
    st = torch.cuda.current_stream().cuda_stream
    for res in (0.01, 0.003):
        ok1 = torch.empty(len(ij), dtype=torch.uint8, device="cuda")
        ok2 = torch.empty(len(ij), dtype=torch.uint8, device="cuda")
        nok = torch.empty(len(nodes), dtype=torch.uint8, device="cuda")
        space.edge_check_indexed_device(d_nodes.data_ptr(), d_ij.data_ptr(), len(ij), res, ok1.data_ptr(), st)
        space.edge_check_indexed_nodes_device(d_nodes.data_ptr(), len(nodes), d_ij.data_ptr(), len(ij), res,
                                              nok.data_ptr(), ok2.data_ptr(), st)
        torch.cuda.synchronize()
        want = oracle.edge_check_indexed(sp, nodes, ij, res)
        assert (ok1.cpu().numpy() == want).all()
        assert (ok2.cpu().numpy() == want).all()
        assert (nok.cpu().numpy() == oracle.feasible(sp, nodes)).all()
