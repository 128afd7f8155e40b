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


# ====================================================================================== spatial sharding
class NumpyShardOps(object):
    """CPU stand-in for sharded._CudaShardOps (same five methods + the mutation hooks): the oracle answers the local
    queries, NumPy does what the routing / finalize / unpack / merge kernels do.  TEST ONLY."""

    class _NN(object):
        def __init__(self, outer):
            self.o = outer

        def add(self, pts):
            self.o.pts = np.concatenate([self.o.pts, np.asarray(pts, dtype=np.float64).reshape(-1, self.o.spec.dim)])
            self.o.mask = np.concatenate([self.o.mask, np.zeros(len(pts), np.uint8)])

        def remove(self, li):
            self.o.mask[li] = 1

        def set_mask(self, m):
            self.o.rej = None if m is None else np.asarray(m, np.uint8)

    def __init__(self, spec):
        self.spec = spec
        self.pts = np.zeros((0, spec.dim))
        self.mask = np.zeros(0, np.uint8)
        self.dead = np.zeros(0, np.uint8)
        self.rej = None
        self.nn = NumpyShardOps._NN(self)

    def _m(self):
        m = self.mask.copy()
        if self.rej is not None:
            m |= self.rej
        return m if m.any() else None

    def set_points(self, pts):
        self.pts = pts.numpy().copy()
        self.mask = np.zeros(len(self.pts), np.uint8)
        self.dead = np.zeros(len(self.pts), np.uint8)

    def route(self, q, axis, bounds, n_ranks):
        qn = q.numpy()
        dest = np.searchsorted(np.asarray(bounds, dtype=np.float64), qn[:, axis], side="right") if n_ranks > 1 \
            else np.zeros(len(qn), np.int64)
        perm = np.argsort(dest, kind="stable")
        counts = np.bincount(dest, minlength=n_ranks).tolist()
        return torch.from_numpy(qn[perm].copy()), torch.from_numpy(perm.astype(np.int64)), counts

    def local_knn(self, q, k):
        if len(q) == 0 or len(self.pts) == 0:
            return (torch.full((len(q), k), -1, dtype=torch.int64), torch.full((len(q), k), np.inf, dtype=torch.float64))
        i, d, _ = oracle.nn_knearest(self.spec, self.pts, q.numpy(), k, self._m())
        return torch.from_numpy(i), torch.from_numpy(d)

    def local_neighbors(self, q, radius, inclusive):
        if len(q) == 0 or len(self.pts) == 0:
            return torch.zeros(len(q) + 1, dtype=torch.int64), torch.zeros(0, dtype=torch.int64)
        off, hits = oracle.nn_neighbors(self.spec, self.pts, q.numpy(), radius, inclusive, self._m())
        return torch.from_numpy(off.astype(np.int64)), torch.from_numpy(hits.astype(np.int64))

    def finalize(self, q, axis, lo_cov, hi_cov, k, idx_local, dist_, l2g, halo, drop_halo, want_flags):
        il, dl = idx_local.numpy(), dist_.numpy()
        g = np.where(il >= 0, l2g.numpy()[np.maximum(il, 0)], -1)
        if drop_halo:
            g = np.where((il >= 0) & (halo.numpy()[np.maximum(il, 0)] != 0), -1, g)
        d = np.where(g >= 0, dl, np.inf)
        rec = np.stack([d.view(np.int64), g], axis=2)
        flags = None
        if want_flags:
            full = (il >= 0).all(1)
            kth = dl[:, k - 1]
            y = q.numpy()[:, axis]
            margin = np.minimum(y - lo_cov, hi_cov - y)
            flags = torch.from_numpy((~(full & (kth < margin))).astype(np.uint8))
        return torch.from_numpy(rec.copy()), flags

    def unpack(self, rec, perm, nq_out, k):
        r = rec.numpy()
        idx = np.full((nq_out, k), -1, np.int64)
        dst = np.full((nq_out, k), np.inf)
        p = perm.numpy()
        idx[p] = r[:, :, 1]
        dst[p] = r[:, :, 0].copy().view(np.float64)
        return torch.from_numpy(idx), torch.from_numpy(dst)

    def merge(self, rec_g, k):
        r = rec_g.numpy()
        return _numpy_merge(torch.from_numpy(r[..., 1].copy()), torch.from_numpy(r[..., 0].copy().view(np.float64)), k)


def _spatial_worker(rank, world, port, D, k, n_total, nq, halo, out_dir):
    from pyoptimalmotionplanning_b200.sharded import SpatialShardedNearestNeighbors
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    pts = rng.random((n_total, D))
    pts[7] = pts[n_total - 2]                       # duplicate across slices: lowest global index wins
    pts[:200, D - 1] *= 0.01                        # a dense strip: unbalanced if the slabs were not quantiles
    q = rng.random((nq, D)) * 1.2 - 0.1             # some outside the bounding box (no certificate -> fallback)
    q[0] = pts[7]
    m = MetricSpec.euclidean(D)
    b, e = stripe(n_total, rank, world)
    s = SpatialShardedNearestNeighbors(m, halo=halo, ops=NumpyShardOps(m))
    s.set(torch.from_numpy(pts[b:e].copy()), b)
    qb, qe = stripe(nq, rank, world)
    myq = torch.from_numpy(q[qb:qe].copy())
    idx, dst, st = s.knearest(myq, k)
    off, hits = s.neighbors(myq, 0.05)              # <= halo when halo is 0.08: routed; else the all-ranks path
    # mutation routed to the holders: add two points (SPMD), remove one old and one new, query again
    g1 = s.add([0.5] * D)
    g2 = s.add([0.25] * (D - 1) + [s.bounds[0] if world > 1 else 0.3])   # exactly on a slab face
    s.remove(7)
    s.remove(g1)
    idx2, dst2, _ = s.knearest(myq, k)
    np.savez(os.path.join(out_dir, "sp%d.npz" % rank), idx=idx.numpy(), dst=dst.numpy(), off=off.numpy(),
             hits=hits.numpy(), idx2=idx2.numpy(), dst2=dst2.numpy(), owned=s.n_owned, unc=st["uncertified"],
             g=np.array([g1, g2]), bounds=np.array(s.bounds))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k,halo", [(2, 1, 0.08), (3, 6, 0.08), (2, 4, 0.004), (1, 3, 0.08)])
def test_spatially_sharded_nn_equals_single_index(tmp_path, world, k, halo):
    D, n_total, nq = 2, 4000, 90 * max(world, 2)
    port = _free_port()
    mp.spawn(_spatial_worker, args=(world, port, D, k, n_total, nq, halo, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    pts = rng.random((n_total, D))
    pts[7] = pts[n_total - 2]
    pts[:200, D - 1] *= 0.01
    q = rng.random((nq, D)) * 1.2 - 0.1
    q[0] = pts[7]
    m = MetricSpec.euclidean(D)
    res = [np.load(tmp_path / ("sp%d.npz" % r)) for r in range(world)]
    idx = np.concatenate([r["idx"] for r in res])
    dst = np.concatenate([r["dst"] for r in res])
    oi, od, _ = oracle.nn_knearest(m, pts, q, k)
    assert (idx == oi).all() and (dst == od).all()
    assert idx[0, 0] == 7
    # balanced shards (quantile boundaries), every point owned exactly once
    owned = [int(r["owned"]) for r in res]
    assert sum(owned) == n_total + 2 and max(owned) - min(owned) <= 0.15 * n_total / world + 2    # + the two added points
    if halo < 0.01 and world > 1:
        assert sum(int(r["unc"]) for r in res) > 0          # the fallback path really ran
    # radius queries
    ooff, ohits = oracle.nn_neighbors(m, pts, q, 0.05, True)
    base = 0
    for r, (qb, qe) in zip(res, [stripe(nq, r, world) for r in range(world)]):
        off, hits = r["off"], r["hits"]
        for j in range(qe - qb):
            want = ohits[ooff[qb + j]:ooff[qb + j + 1]]
            assert (hits[off[j]:off[j + 1]] == want).all()
    # after add / remove
    g1, g2 = int(res[0]["g"][0]), int(res[0]["g"][1])
    assert (g1, g2) == (n_total, n_total + 1)
    b0 = float(res[0]["bounds"][0]) if world > 1 else 0.3
    pts2 = np.concatenate([pts, [[0.5] * D], [[0.25] * (D - 1) + [b0]]])
    mask = np.zeros(len(pts2), np.uint8)
    mask[[7, g1]] = 1
    oi2, od2, _ = oracle.nn_knearest(m, pts2, q, k, mask)
    idx2 = np.concatenate([r["idx2"] for r in res])
    dst2 = np.concatenate([r["dst2"] for r in res])
    assert (idx2 == oi2).all() and (dst2 == od2).all()
