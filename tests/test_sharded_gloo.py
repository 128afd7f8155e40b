"""The N>1 path on CPU: world_size-2/3 gloo process groups run the same all-gather -> local top-k
-> all-to-all -> merge pipeline as the NCCL path, with oracle-backed local search and a numpy
merge standing in for the two CUDA kernels.  The answer must equal the single-shard answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from pyoptimalmotionplanning_b200.descriptors import MetricSpec
from pyoptimalmotionplanning_b200.sharded import ShardedNearestNeighbors, stripe


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _numpy_merge(idx_in, dist_in, k):
    G, nq, _ = idx_in.shape
    ii = idx_in.permute(1, 0, 2).reshape(nq, G * k).numpy()
    dd = dist_in.permute(1, 0, 2).reshape(nq, G * k).numpy().copy()
    dd[ii < 0] = np.inf
    key_i = np.where(ii < 0, np.iinfo(np.int64).max, ii)
    order = np.lexsort((key_i, dd), axis=1)[:, :k]
    oi = np.take_along_axis(ii, order, 1)
    od = np.take_along_axis(dd, order, 1)
    oi[np.isinf(od)] = -1
    return torch.from_numpy(oi), torch.from_numpy(od)


def _worker(rank, world, port, D, k, n_total, nq, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(77)
    pts = rng.random((n_total, D))
    pts[5] = pts[n_total - 3]                      # a cross-shard duplicate: tie broken by global index
    q = rng.random((nq, D))
    q[0] = pts[5]
    b, e = stripe(n_total, rank, world)
    m = MetricSpec.euclidean(D)
    shard = pts[b:e]

    def local_knn(q_all, kk):
        i, d, _ = oracle.nn_knearest(m, shard, q_all.numpy(), kk)
        return torch.from_numpy(i), torch.from_numpy(d)

    s = ShardedNearestNeighbors(b, local_knn=local_knn, merge=_numpy_merge)
    qs = nq // world
    idx, dst = s.knearest(torch.from_numpy(q[rank * qs:(rank + 1) * qs].copy()), k)
    np.save(os.path.join(out_dir, "idx%d.npy" % rank), idx.numpy())
    np.save(os.path.join(out_dir, "dst%d.npy" % rank), dst.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 1), (2, 5), (3, 4)])
def test_sharded_knn_equals_single_shard(tmp_path, world, k):
    D, n_total, nq = 3, 3001, 60 * world
    port = _free_port()
    mp.spawn(_worker, args=(world, port, D, k, n_total, nq, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(77)
    pts = rng.random((n_total, D))
    pts[5] = pts[n_total - 3]
    q = rng.random((nq, D))
    q[0] = pts[5]
    oi, od, _ = oracle.nn_knearest(MetricSpec.euclidean(D), pts, q, k)
    idx = np.concatenate([np.load(tmp_path / ("idx%d.npy" % r)) for r in range(world)])
    dst = np.concatenate([np.load(tmp_path / ("dst%d.npy" % r)) for r in range(world)])
    assert (idx == oi).all() and (dst == od).all()
    assert idx[0, 0] == 5                           # lowest global index wins the exact tie


def test_stripe_covers_everything():
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 3, 8):
            spans = [stripe(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _stripe_worker(rank, world, port, out_dir):
    """Edges and lock-step planner trials are independent units: every rank takes its stripe, no data-path
    collective; the only communication is the final gather of the results (here: files + a barrier)."""
    import random
    from pyoptimalmotionplanning_b200 import SpaceSpec
    from pyoptimalmotionplanning_b200.rrt import RepeatedRRTForest
    from tests.fake_backend import OracleForest, OracleSpace
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp = SpaceSpec([0, 0], [1, 1], boxes=[[0.3, 0.3, 0.7, 0.7]])
    rng = np.random.default_rng(3)
    a, b = rng.random((500, 2)), rng.random((500, 2))
    e0, e1 = stripe(500, rank, world)
    ok = oracle.edge_check(sp, a[e0:e1], b[e0:e1], 0.01)
    flags = [None] * world
    dist.all_gather_object(flags, (e0, np.asarray(ok, bool)))
    if rank == 0:
        full = np.concatenate([f[1] for f in sorted(flags, key=lambda t: t[0])])
        np.save(os.path.join(out_dir, "edges.npy"), full)
    t0, t1 = stripe(5, rank, world)
    rf = RepeatedRRTForest(sp, [0.1, 0.5], list(range(t0, t1)), [0.9, 0.5], 0.1, max_step=0.15, resolution=0.01,
                           _backends=(OracleForest, OracleSpace))
    for _ in range(30):
        rf.planMore(10)
    costs = [None] * world
    dist.all_gather_object(costs, (t0, rf.bestPathCost))
    if rank == 0:
        np.save(os.path.join(out_dir, "costs.npy"),
                np.array([c if c is not None else np.inf for _, cs in sorted(costs) for c in cs]))
    dist.barrier()
    dist.destroy_process_group()


def test_striped_edges_and_trials_equal_single_process(tmp_path):
    from pyoptimalmotionplanning_b200 import SpaceSpec
    from pyoptimalmotionplanning_b200.rrt import RepeatedRRTForest
    from tests.fake_backend import OracleForest, OracleSpace
    port = _free_port()
    mp.spawn(_stripe_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sp = SpaceSpec([0, 0], [1, 1], boxes=[[0.3, 0.3, 0.7, 0.7]])
    rng = np.random.default_rng(3)
    a, b = rng.random((500, 2)), rng.random((500, 2))
    assert (np.load(tmp_path / "edges.npy") == np.asarray(oracle.edge_check(sp, a, b, 0.01), bool)).all()
    rf = RepeatedRRTForest(sp, [0.1, 0.5], list(range(5)), [0.9, 0.5], 0.1, max_step=0.15, resolution=0.01,
                           _backends=(OracleForest, OracleSpace))
    for _ in range(30):
        rf.planMore(10)
    want = np.array([c if c is not None else np.inf for c in rf.bestPathCost])
    assert (np.load(tmp_path / "costs.npy") == want).all()      # trial t depends on seed t only, not on the striping
