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
