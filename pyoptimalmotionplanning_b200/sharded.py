"""Multi-GPU nearest neighbours: the node set sharded across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Rank r owns a contiguous block
of insertion indices [offset_r, offset_r + n_r).  A k-NN batch is answered in four steps:

  1. all-gather the per-rank query slices so every rank sees the whole batch        (NCCL)
  2. local exact top-k of every query against the local shard                       (grid kernel)
  3. all-to-all: rank r receives, for ITS query slice, the candidates of every shard (NCCL)
  4. merge the G candidate lists per query by (distance, global index)               (merge kernel)

The merge order makes the result independent of the number of shards: it equals the 1-GPU answer
bit for bit.  Edge checks and propagation batches need no collective: their units are independent
and are simply striped across ranks.

`local_knn` / `merge` are injectable so the collective plumbing can be exercised with the gloo
backend on CPU (tests/test_sharded_gloo.py supplies oracle-backed functions); the defaults are the
CUDA kernels.
"""
import torch
import torch.distributed as dist


def _cuda_local_knn(nn):
    def fn(q_all, k):
        nq = q_all.shape[0]
        idx = torch.empty((nq, k), dtype=torch.int64, device=q_all.device)
        dst = torch.empty((nq, k), dtype=torch.float64, device=q_all.device)
        nn.knearest_device(q_all.data_ptr(), nq, k, idx.data_ptr(), dst.data_ptr(), None,
                           torch.cuda.current_stream().cuda_stream)
        return idx, dst
    return fn


def _cuda_merge(idx_in, dist_in, k):
    from .backend import merge_candidates_device
    G, nq = idx_in.shape[0], idx_in.shape[1]
    idx = torch.empty((nq, k), dtype=torch.int64, device=idx_in.device)
    dst = torch.empty((nq, k), dtype=torch.float64, device=idx_in.device)
    merge_candidates_device(idx_in.data_ptr(), dist_in.data_ptr(), G, nq, k, idx.data_ptr(), dst.data_ptr(), None,
                            torch.cuda.current_stream().cuda_stream)
    return idx, dst


class ShardedNearestNeighbors(object):
    def __init__(self, global_offset, local_knn=None, merge=None, nn=None, group=None):
        """global_offset: insertion index of this rank's first point.
        nn: a backend.CudaNN holding the local shard (used when local_knn is not given)."""
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.offset = int(global_offset)
        self.local_knn = local_knn or _cuda_local_knn(nn)
        self.merge = merge or _cuda_merge

    def knearest(self, q_slice, k):
        """q_slice: [Qs, D] float64 tensor (same Qs on every rank).  Returns (idx[Qs,k] global int64,
        dist[Qs,k]) for this rank's slice."""
        G = self.world
        Qs, D = q_slice.shape
        q_all = torch.empty((G * Qs, D), dtype=q_slice.dtype, device=q_slice.device)
        dist.all_gather_into_tensor(q_all, q_slice.contiguous(), group=self.group)
        idx_l, dst_l = self.local_knn(q_all, k)
        idx_l = torch.where(idx_l >= 0, idx_l + self.offset, idx_l)  # local -> global insertion index
        idx_r = torch.empty_like(idx_l)
        dst_r = torch.empty_like(dst_l)
        # row block g of the send buffer goes to rank g; block g of the receive buffer came from rank g
        dist.all_to_all_single(idx_r, idx_l.contiguous(), group=self.group)
        dist.all_to_all_single(dst_r, dst_l.contiguous(), group=self.group)
        return self.merge(idx_r.view(G, Qs, k), dst_r.view(G, Qs, k), k)

    def nearest(self, q_slice):
        idx, dst = self.knearest(q_slice, 1)
        return idx[:, 0], dst[:, 0]


def stripe(n_units, rank, world):
    """Contiguous share [begin, end) of n independent units (edges, (x,u) pairs, trials)."""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
