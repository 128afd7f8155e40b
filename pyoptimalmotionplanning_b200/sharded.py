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


# ====================================================================================== spatial sharding
class _CudaShardOps(object):
    """Device side of SpatialShardedNearestNeighbors: the local index (backend.CudaNN) and the routing / finalize /
    unpack / merge kernels of csrc/shard.cu and csrc/nn.cu, on torch CUDA tensors.  tests/test_sharded_gloo.py
    substitutes a NumPy/oracle implementation of the same five methods to run the collective plumbing on CPU."""

    def __init__(self, metric_spec, device):
        import ctypes as C
        from . import _abi
        from .backend import CudaNN
        self.C, self._abi = C, _abi
        self.lib = _abi.load_library()
        self.spec = metric_spec
        self.device = torch.device("cuda", device)
        self.nn = CudaNN(metric_spec, device)
        self._scratch = None

    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def set_points(self, pts):
        self.pts = pts.contiguous()
        if pts.shape[0]:
            self.nn.set_device(self.pts.data_ptr(), pts.shape[0], self._st())
            self.nn.build_index(self._st())
        else:
            self.nn.reset()

    def route(self, q, axis, bounds, n_ranks):
        C = self.C
        nq, D = q.shape
        dest = torch.empty(max(nq, 1), dtype=torch.int32, device=q.device)
        cdev = torch.empty(128, dtype=torch.int64, device=q.device)
        counts = (C.c_int64 * n_ranks)()
        b = (C.c_double * max(n_ranks - 1, 1))(*[float(x) for x in bounds])
        q_out = torch.empty_like(q)
        perm = torch.empty(nq, dtype=torch.int64, device=q.device)
        self._abi.check(self.lib, self.lib.pomp_shard_route(
            C.c_void_p(q.data_ptr()), nq, D, axis, b, n_ranks, C.c_void_p(dest.data_ptr()), C.c_void_p(cdev.data_ptr()),
            counts, C.c_void_p(q_out.data_ptr()), C.c_void_p(perm.data_ptr()), C.c_void_p(self._st())))
        return q_out, perm, [int(c) for c in counts]

    def local_knn(self, q, k):
        nq = q.shape[0]
        if nq and self.pts.shape[0]:
            # every row is written by the search (missing neighbours as -1 / inf): no fill kernels
            idx = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            dst = torch.empty((nq, k), dtype=torch.float64, device=q.device)
            self.nn.knearest_device(q.data_ptr(), nq, k, idx.data_ptr(), dst.data_ptr(), None, self._st())
            return idx, dst
        idx = torch.full((nq, k), -1, dtype=torch.int64, device=q.device)
        dst = torch.full((nq, k), float("inf"), dtype=torch.float64, device=q.device)
        return idx, dst

    def local_neighbors(self, q, radius, inclusive):
        nq = q.shape[0]
        off = torch.zeros(nq + 1, dtype=torch.int64, device=q.device)
        if nq == 0 or self.pts.shape[0] == 0:
            return off, torch.empty(0, dtype=torch.int64, device=q.device)
        cap = max(64, 32 * nq)
        while True:
            hits = torch.empty(cap, dtype=torch.int64, device=q.device)
            rc, need = self.nn.neighbors_device(q.data_ptr(), nq, radius, inclusive, off.data_ptr(), hits.data_ptr(), cap,
                                                self._st())
            if rc == self._abi.POMP_ERR_CAPACITY:
                cap = int(need)
                continue
            return off, hits[:need]

    def finalize(self, q, axis, lo_cov, hi_cov, k, idx_local, dist, l2g, halo, drop_halo, want_flags):
        C = self.C
        nq, D = q.shape
        rec = torch.empty((nq, k, 2), dtype=torch.int64, device=q.device)
        flags = torch.zeros(nq, dtype=torch.uint8, device=q.device) if want_flags else None
        self._abi.check(self.lib, self.lib.pomp_shard_finalize(
            C.c_void_p(q.data_ptr()), nq, D, axis, float(lo_cov), float(hi_cov), k, C.c_void_p(idx_local.data_ptr()),
            C.c_void_p(dist.data_ptr()), C.c_void_p(l2g.data_ptr()), C.c_void_p(halo.data_ptr()), 1 if drop_halo else 0,
            C.c_void_p(rec.data_ptr()), C.c_void_p(flags.data_ptr() if want_flags else 0), C.c_void_p(self._st())))
        return rec, flags

    def unpack(self, rec, perm, nq_out, k):
        C = self.C
        n = rec.shape[0]
        idx = torch.full((nq_out, k), -1, dtype=torch.int64, device=rec.device)
        dst = torch.full((nq_out, k), float("inf"), dtype=torch.float64, device=rec.device)
        self._abi.check(self.lib, self.lib.pomp_shard_unpack(
            C.c_void_p(rec.data_ptr()), n, k, C.c_void_p(perm.data_ptr() if perm is not None else 0),
            C.c_void_p(idx.data_ptr()), C.c_void_p(dst.data_ptr()), None, C.c_void_p(self._st())))
        return idx, dst

    def merge(self, rec_g, k):
        """rec_g: [G, m, k, 2] records -> (idx[m,k], dist[m,k]) by (distance, global index)."""
        G, m = rec_g.shape[0], rec_g.shape[1]
        idx_in = rec_g[..., 1].contiguous()
        dist_in = rec_g[..., 0].contiguous().view(torch.float64)
        return _cuda_merge(idx_in, dist_in, k)


class _PendingKnn(object):
    """A sharded k-NN batch whose kernels are in flight (SpatialShardedNearestNeighbors.knearest_p2p_async)."""

    def __init__(self, owner, q, k, idx, dst, dlist, dcount, pin, event, cap, fallback_slots, mark):
        self.owner, self.q, self.k, self.idx, self.dst = owner, q, k, idx, dst
        self.dlist, self.dcount, self.pin, self.event = dlist, dcount, pin, event
        self.cap, self.fallback_slots, self._mark = cap, fallback_slots, mark

    def wait(self):
        """Collective (same order on every rank).  (idx[nq,k] global int64, dist[nq,k], stats)."""
        self.event.synchronize()
        n_def, n_worst = int(self.pin[0]), int(self.pin[1])
        self._mark("host read")
        if n_worst > 0:
            self.owner._fallback(self.q, self.k, self.dlist, self.dcount, n_worst, self.fallback_slots, self.idx, self.dst)
            self._mark("fallback")
        return self.idx, self.dst, {"deferred": n_def, "deferred_worst_rank": n_worst, "block_slots": self.cap}


class SpatialShardedNearestNeighbors(object):
    """The node set of one NearestNeighbors sharded SPATIALLY over the ranks of a process group.

    Rank r owns the points whose coordinate `axis` falls into its slab [b_r, b_{r+1}) -- boundaries are global
    quantiles, so the shards are balanced -- and keeps HALO copies of its neighbours' points within `halo` of the slab
    faces.  A k-NN batch is answered with two all-to-alls and no replicated work:

      1. route   each query goes to the rank owning its slab                      (counting sort + all_to_all)
      2. search  local exact k-NN over owned + halo points                         (grid kernels)
      3. certify k-th distance < distance to the outside of the covered region     (finalize kernel)
      4. return  packed 16-byte (distance, GLOBAL index) records to the source     (all_to_all)

    Queries without a certificate (sparse regions, k large against the halo) take the exact fallback: they are
    all-gathered, every rank answers them over its OWNED points, the G candidate lists are merged by (distance,
    global index).  Either way the answer equals the one-GPU answer over the union bit for bit.  neighbors(radius)
    follows the same route (certified when radius <= halo).  add / remove / set_mask go to the owning rank (and to
    the neighbour holding a halo copy).  Global insertion indices are the caller's: set(points, first_index).
    """

    def __init__(self, metric_spec, axis=None, halo=None, group=None, device=None, ops=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.spec = metric_spec
        self.dim = metric_spec.dim
        self.axis = self.dim - 1 if axis is None else int(axis)
        self.halo_arg = halo
        self.ops = ops if ops is not None else _CudaShardOps(metric_spec, device if device is not None else 0)
        self.bounds = []
        self.n_global = 0
        self.inclusive_radius = True

    # ------------------------------------------------------------------ collectives with ragged splits
    def _a2a(self, send, send_counts, row_shape=()):
        """all_to_all of rows grouped by destination rank; returns (received rows, recv_counts)."""
        G = self.world
        sc = torch.tensor(send_counts, dtype=torch.int64, device=send.device)
        rc = torch.empty(G, dtype=torch.int64, device=send.device)
        dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = [int(c) for c in rc.tolist()]
        recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=recv_counts, input_split_sizes=list(send_counts),
                               group=self.group)
        return recv, recv_counts

    def _a2a_back(self, send, send_counts, recv_counts):
        recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=list(recv_counts),
                               input_split_sizes=list(send_counts), group=self.group)
        return recv

    # ------------------------------------------------------------------ build
    def set(self, pts, first_index):
        """Collective.  pts: this rank's [n_r, D] float64 slice of the global node set (any split), whose rows carry
        the global insertion indices first_index .. first_index + n_r - 1."""
        G, ax = self.world, self.axis
        dev = pts.device
        n_r = pts.shape[0]
        tot = torch.tensor([n_r], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=self.group)
        self.n_global = int(tot.item())
        # slab boundaries: global quantiles of a fixed-size sample of the axis coordinate
        S = 4096
        if n_r:
            pick = torch.linspace(0, n_r - 1, S, device=dev).long()
            samp = torch.sort(pts[:, ax])[0][pick]
        else:
            samp = torch.full((S,), float("nan"), dtype=torch.float64, device=dev)
        allsamp = torch.empty(G * S, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allsamp, samp.contiguous(), group=self.group)
        allsamp = torch.sort(allsamp[~torch.isnan(allsamp)])[0]
        m = allsamp.numel()
        self.bounds = [float(allsamp[min(m - 1, (m * r) // G)].item()) for r in range(1, G)] if m else [0.0] * (G - 1)
        lo_all = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
        hi_all = torch.full((1,), float("-inf"), dtype=torch.float64, device=dev)
        # halo width: a ball holding ~(32 + 8 k_ref) points at the mean density (k_ref = 16), or the caller's value
        ext_lo = pts.min(0)[0] if n_r else torch.full((self.dim,), float("inf"), dtype=torch.float64, device=dev)
        ext_hi = pts.max(0)[0] if n_r else torch.full((self.dim,), float("-inf"), dtype=torch.float64, device=dev)
        dist.all_reduce(ext_lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(ext_hi, op=dist.ReduceOp.MAX, group=self.group)
        del lo_all, hi_all
        if self.halo_arg is not None:
            self.halo = float(self.halo_arg)
        else:
            import math
            D = self.dim
            vol = float(torch.clamp(ext_hi - ext_lo, min=1e-300).prod().item())
            vball = math.pi ** (D / 2.0) / math.gamma(D / 2.0 + 1.0)
            self.halo = (160.0 * vol / (max(self.n_global, 1) * vball)) ** (1.0 / D)
        b = torch.tensor(self.bounds, dtype=torch.float64, device=dev)
        y = pts[:, ax].contiguous()
        owner = torch.bucketize(y, b, right=True) if G > 1 else torch.zeros(n_r, dtype=torch.int64, device=dev)
        gidx = torch.arange(first_index, first_index + n_r, dtype=torch.int64, device=dev)
        # rows to send: (dest, gidx, halo flag) -- the owner's copy, plus halo copies for the adjacent slabs
        dests, rows, flags = [owner], [torch.arange(n_r, device=dev)], [torch.zeros(n_r, dtype=torch.uint8, device=dev)]
        for r in range(G):
            if r > 0:      # rank r's lower face b[r-1]: points just below it belong to r-1 and are halo of r
                msk = (owner < r) & (y >= self.bounds[r - 1] - self.halo)
                sel = torch.nonzero(msk).flatten()
                dests.append(torch.full_like(sel, r))
                rows.append(sel)
                flags.append(torch.ones(sel.numel(), dtype=torch.uint8, device=dev))
            if r < G - 1:  # upper face b[r]
                msk = (owner > r) & (y <= self.bounds[r] + self.halo)
                sel = torch.nonzero(msk).flatten()
                dests.append(torch.full_like(sel, r))
                rows.append(sel)
                flags.append(torch.ones(sel.numel(), dtype=torch.uint8, device=dev))
        dest = torch.cat(dests)
        row = torch.cat(rows)
        flag = torch.cat(flags)
        order = torch.sort(dest, stable=True)[1]
        dest, row, flag = dest[order], row[order], flag[order]
        counts = torch.bincount(dest, minlength=G).tolist()
        payload = torch.cat([pts[row], gidx[row].view(torch.float64).unsqueeze(1),
                             flag.to(torch.float64).unsqueeze(1)], 1)
        recv, _ = self._a2a(payload, counts)
        g_in = recv[:, self.dim].contiguous().view(torch.int64)
        order = torch.sort(g_in)[1]          # local order = ascending global index (global -> local by binary search)
        self.l2g = g_in[order].contiguous()
        self.halo_flag = recv[order, self.dim + 1].to(torch.uint8).contiguous()
        self.local_pts = recv[order, :self.dim].contiguous()
        self.ops.set_points(self.local_pts)
        self.lo_cov = float("-inf") if self.rank == 0 else self.bounds[self.rank - 1] - self.halo
        self.hi_cov = float("inf") if self.rank == G - 1 else self.bounds[self.rank] + self.halo
        self.n_owned = int((self.halo_flag == 0).sum().item())
        # one ordinary point inside every slab: padding of the fixed-size exchange blocks (cheap to answer)
        mid = 0.5 * (ext_lo + ext_hi)
        edges = [float(ext_lo[ax].item())] + list(self.bounds) + [float(ext_hi[ax].item())]
        self.pad_pts = mid.repeat(G, 1).contiguous()
        for r in range(G):
            self.pad_pts[r, ax] = 0.5 * (edges[r] + edges[r + 1])
        return self

    # ------------------------------------------------------------------ queries
    def knearest(self, q, k):
        """Collective.  q: this rank's [nq_r, D] batch (sizes may differ between ranks).  Returns (idx[nq_r,k] global
        int64, dist[nq_r,k], stats dict)."""
        G, ax = self.world, self.axis
        nq = q.shape[0]
        qs, perm, counts = self.ops.route(q.contiguous(), ax, self.bounds, G)
        qr, rcounts = self._a2a(qs, counts)
        il, dl = self.ops.local_knn(qr, k)
        rec, flags = self.ops.finalize(qr, ax, self.lo_cov, self.hi_cov, k, il, dl, self.l2g, self.halo_flag, False, True)
        back = self._a2a_back(torch.cat([rec.view(rec.shape[0], -1), flags.to(torch.int64).unsqueeze(1)], 1), rcounts, counts)
        rec_b = back[:, :2 * k].contiguous().view(-1, k, 2)
        unc_sorted = back[:, 2 * k] != 0
        idx, dst = self.ops.unpack(rec_b, perm, nq, k)
        # exact fallback for the queries without a certificate (collective decision: any rank has one)
        unc_pos = perm[unc_sorted]
        n_unc = torch.tensor([unc_pos.numel()], dtype=torch.int64, device=q.device)
        mx = n_unc.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        m = int(mx.item())
        if m > 0:
            qf = torch.zeros((m, self.dim), dtype=torch.float64, device=q.device)
            qf[:unc_pos.numel()] = q[unc_pos]
            allq = torch.empty((G * m, self.dim), dtype=torch.float64, device=q.device)
            dist.all_gather_into_tensor(allq, qf, group=self.group)
            ia, da = self.ops.local_knn(allq, k)
            reca, _ = self.ops.finalize(allq, ax, self.lo_cov, self.hi_cov, k, ia, da, self.l2g, self.halo_flag, True, False)
            got = torch.empty_like(reca)                       # block g: rank g's candidates for MY m queries
            dist.all_to_all_single(got, reca.contiguous(), group=self.group)
            mi, md = self.ops.merge(got.view(G, m, k, 2), k)
            nu = unc_pos.numel()
            idx[unc_pos] = mi[:nu]
            dst[unc_pos] = md[:nu]
        return idx, dst, {"routed": counts, "uncertified": int(n_unc.item())}

    def nearest(self, q):
        idx, dst, st = self.knearest(q, 1)
        return idx[:, 0], dst[:, 0], st

    def knearest_fast(self, q, k, slack=1.08, fallback_slots=4096, _trace=None):
        """The same answer as knearest() for batches of EQUAL size on every rank, without a host synchronisation on
        the way: fixed-capacity blocks (nq / G * slack + 1024 slots per destination) make both all-to-alls equal-split.
        ONE synchronisation at the end reads the number of deferred queries (no certificate, or no room in a block);
        only if there are any does the exact all-ranks fallback run, over a block sized for the worst rank.  CUDA
        only.  Returns (idx, dist, stats)."""
        ops, G, ax, D = self.ops, self.world, self.axis, self.dim
        C, lib, check, st = ops.C, ops.lib, ops._abi.check, ops._st()

        def mark(name):   # optional stage trace: (name, CUDA event, host clock)
            if _trace is not None:
                import time as _t
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                _trace.append((name, e, _t.perf_counter()))
        mark("start")
        nq = q.shape[0]
        dev = q.device
        q = q.contiguous()
        cap = int(nq / G * slack) + 1024
        n_slots = G * cap
        # unused slots: an ordinary point of the destination slab (cheap to answer) / position -1
        sendq = self.pad_pts.view(G, 1, D).expand(G, cap, D).contiguous().view(n_slots, D)
        perm = torch.full((n_slots,), -1, dtype=torch.int64, device=dev)
        ctr = torch.zeros(72, dtype=torch.int64, device=dev)
        dlist = torch.empty(max(nq, 1), dtype=torch.int64, device=dev)
        b = (C.c_double * max(G - 1, 1))(*[float(x) for x in self.bounds])
        check(lib, lib.pomp_shard_route_fixed(C.c_void_p(q.data_ptr()), nq, D, ax, b, G, cap, C.c_void_p(ctr.data_ptr()),
                                              C.c_void_p(sendq.data_ptr()), C.c_void_p(perm.data_ptr()),
                                              C.c_void_p(dlist.data_ptr()), nq, C.c_void_p(st)))
        mark("route")
        recvq = torch.empty_like(sendq)
        dist.all_to_all_single(recvq, sendq, group=self.group)
        mark("a2a_queries")
        il, dl = ops.local_knn(recvq, k)
        mark("local_search")
        rows = torch.empty((n_slots, 2 * k + 1), dtype=torch.int64, device=dev)
        check(lib, lib.pomp_shard_finalize_rows(C.c_void_p(recvq.data_ptr()), n_slots, D, ax, float(self.lo_cov),
                                                float(self.hi_cov), k, C.c_void_p(il.data_ptr()), C.c_void_p(dl.data_ptr()),
                                                C.c_void_p(self.l2g.data_ptr()), C.c_void_p(self.halo_flag.data_ptr()), 0,
                                                C.c_void_p(rows.data_ptr()), C.c_void_p(st)))
        mark("finalize")
        back = torch.empty_like(rows)
        dist.all_to_all_single(back, rows, group=self.group)
        mark("a2a_results")
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dst = torch.empty((nq, k), dtype=torch.float64, device=dev)
        dcount = ctr[64:65]
        check(lib, lib.pomp_shard_unpack_rows(C.c_void_p(back.data_ptr()), n_slots, k, C.c_void_p(perm.data_ptr()),
                                              C.c_void_p(idx.data_ptr()), C.c_void_p(dst.data_ptr()),
                                              C.c_void_p(dlist.data_ptr()), C.c_void_p(dcount.data_ptr()), nq,
                                              C.c_void_p(st)))
        mark("unpack")
        # how many queries were deferred, anywhere?  (the one synchronisation of the step: it also tells the caller
        # that the answers are complete)  None, the common case: done.
        worst = dcount.clone()
        if G > 1:
            dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=self.group)
        mark("deferred-count all_reduce")
        n_def, n_worst = self._read_counts(dcount, worst)
        mark("host read")
        if n_worst > 0:
            self._fallback(q, k, dlist, dcount, n_worst, fallback_slots, idx, dst)
            mark("fallback")
        return idx, dst, {"deferred": n_def, "deferred_worst_rank": n_worst, "block_slots": cap}

    # ------------------------------------------------------------------ peer-memory exchange (NVLink stores, no NCCL)
    def _read_counts(self, dcount, worst):
        """(own deferred, worst rank's deferred) on the host: two 8-byte copies into pinned memory and one stream
        synchronisation (torch.cat + tolist costs a kernel and a pageable copy more)."""
        if getattr(self, "_pin2", None) is None:
            self._pin2 = torch.empty(2, dtype=torch.int64).pin_memory()
        self._pin2[0:1].copy_(dcount, non_blocking=True)
        self._pin2[1:2].copy_(worst, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(self._pin2[0]), int(self._pin2[1])

    def _p2p_setup(self, cap, k, dev):
        import torch.distributed._symmetric_memory as symm
        key = (cap, k)
        if getattr(self, "_p2p_key", None) == key:
            return
        G, D, C = self.world, self.dim, self.ops.C
        group = self.group if self.group is not None else dist.group.WORLD
        self._rq = symm.empty(G * cap * D, dtype=torch.float64, device=dev)
        self._rq_h = symm.rendezvous(self._rq, group)
        self._rb = symm.empty(G * cap * (2 * k + 1), dtype=torch.int64, device=dev)
        self._rb_h = symm.rendezvous(self._rb, group)
        # slots that are not written in a step keep older queries: start from ordinary points of this slab
        self._rq.view(G * cap, D)[:] = self.pad_pts[self.rank]
        self._rb.zero_()
        self._rq_ptrs = (C.c_uint64 * G)(*[int(p) for p in self._rq_h.buffer_ptrs])
        self._rb_ptrs = (C.c_uint64 * G)(*[int(p) for p in self._rb_h.buffer_ptrs])
        # deferred counts of a step, replicated on every rank by peer stores: [parity][G owners + 1 (route overflow)][G askers]
        self._fl = symm.empty(2 * (G + 1) * G, dtype=torch.int64, device=dev)
        self._fl_h = symm.rendezvous(self._fl, group)
        self._fl.zero_()
        self._fl_ptrs = (C.c_uint64 * G)(*[int(p) for p in self._fl_h.buffer_ptrs])
        self._p2p_step = 0
        self._bounds_c = (C.c_double * max(G - 1, 1))(*[float(x) for x in self.bounds])
        self._perm = torch.empty(G * cap, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        self._rq_h.barrier(channel=0)
        self._p2p_key = key

    def knearest_p2p(self, q, k, slack=1.04, fallback_slots=4096, _trace=None):
        """knearest_fast() with the two all-to-alls replaced by NVLink stores from the routing / finalize kernels into
        the peers' receive buffers (torch symmetric memory supplies the mapped peer pointers and the device barrier):
        route+send is ONE kernel, finalize+return is ONE kernel, two barriers per step, no NCCL call in the step (the
        deferred counts travel as peer stores ahead of the second barrier).  Same answers."""
        return self.knearest_p2p_async(q, k, slack, fallback_slots, _trace).wait()

    def knearest_p2p_async(self, q, k, slack=1.04, fallback_slots=4096, _trace=None):
        """The same exchange, issued without the host synchronisation at its end.  Returns a pending result whose
        wait() reads the deferred count of THIS batch (a CUDA event recorded behind its two 8-byte copies, not a
        device synchronisation), runs the exact fallback if anything was deferred and hands back (idx, dist, stats).
        A caller that streams batches calls wait() of batch i after issuing batch i + 1: the host read no longer
        idles the GPU.  Collective: every rank issues and waits in the same order; `q` must stay unchanged until
        wait() returns (the fallback re-reads the deferred queries from it)."""
        ops, G, ax, D = self.ops, self.world, self.axis, self.dim
        C, lib, check, st = ops.C, ops.lib, ops._abi.check, ops._st()

        def mark(name):
            if _trace is not None:
                import time as _t
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                _trace.append((name, e, _t.perf_counter()))
        mark("start")
        nq = q.shape[0]
        dev = q.device
        q = q.contiguous()
        cap = int(nq / G * slack) + 1024
        n_slots = G * cap
        self._p2p_setup(cap, k, dev)
        ctr = torch.zeros(72 + 64, dtype=torch.int64, device=dev)   # [0, G) cursors, [64] deferred, [72, 72 + G) fails
        dlist = torch.empty(max(nq, 1), dtype=torch.int64, device=dev)
        b = self._bounds_c
        par_off = (self._p2p_step & 1) * (G + 1) * G
        self._p2p_step += 1
        check(lib, lib.pomp_shard_route_p2p(C.c_void_p(q.data_ptr()), nq, D, ax, b, G, self.rank, cap,
                                            C.c_void_p(ctr.data_ptr()), self._rq_ptrs, C.c_void_p(self._perm.data_ptr()),
                                            C.c_void_p(dlist.data_ptr()), nq, C.c_void_p(st)))
        # my route overflow -> row G, column `rank` of every rank's flag block
        check(lib, lib.pomp_shard_share_counts_p2p(C.c_void_p(ctr.data_ptr() + 64 * 8), 1, self._fl_ptrs,
                                                   par_off + G * G + self.rank, G, C.c_void_p(st)))
        mark("route+send kernel")
        self._rq_h.barrier(channel=0)
        mark("barrier")
        recvq = self._rq.view(n_slots, D)
        il, dl = ops.local_knn(recvq, k)
        mark("local_search")
        check(lib, lib.pomp_shard_finalize_p2p(C.c_void_p(recvq.data_ptr()), cap, G, self.rank, D, ax, float(self.lo_cov),
                                               float(self.hi_cov), k, C.c_void_p(il.data_ptr()), C.c_void_p(dl.data_ptr()),
                                               C.c_void_p(self.l2g.data_ptr()), self._rb_ptrs,
                                               C.c_void_p(ctr.data_ptr() + 72 * 8), C.c_void_p(st)))
        # the queries I could not certify, per asking rank -> row `rank` of every rank's flag block
        check(lib, lib.pomp_shard_share_counts_p2p(C.c_void_p(ctr.data_ptr() + 72 * 8), G, self._fl_ptrs,
                                                   par_off + self.rank * G, G, C.c_void_p(st)))
        mark("finalize+return kernel")
        self._rb_h.barrier(channel=1)
        mark("barrier")
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        dst = torch.empty((nq, k), dtype=torch.float64, device=dev)
        dcount = ctr[64:65]
        check(lib, lib.pomp_shard_unpack_p2p(C.c_void_p(self._rb.data_ptr()), G, cap, k, C.c_void_p(self._perm.data_ptr()),
                                             C.c_void_p(ctr.data_ptr()), C.c_void_p(idx.data_ptr()),
                                             C.c_void_p(dst.data_ptr()), C.c_void_p(dlist.data_ptr()), nq, C.c_void_p(st)))
        mark("unpack")
        # the largest deferred count of any rank, from the flag block every rank now holds: no collective.  (An upper
        # bound: an owner also counts stale padding slots it could not certify; that can only trigger a fallback that
        # finds nothing to do.)
        worst = torch.empty(1, dtype=torch.int64, device=dev)
        check(lib, lib.pomp_shard_worst_count(C.c_void_p(self._fl.data_ptr() + par_off * 8), G + 1, G,
                                              C.c_void_p(worst.data_ptr()), C.c_void_p(st)))
        mark("deferred-count reduce")
        # (own, worst rank's) deferred counts into one of a few rotating pinned buffers, an event behind the copies
        pool = getattr(self, "_pin_pool", None)
        if pool is None:
            pool = self._pin_pool = [torch.empty(2, dtype=torch.int64).pin_memory() for _ in range(8)]
            self._pin_next = 0
        pin = pool[self._pin_next]
        self._pin_next = (self._pin_next + 1) % len(pool)
        pin[0:1].copy_(dcount, non_blocking=True)
        pin[1:2].copy_(worst, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return _PendingKnn(self, q, k, idx, dst, dlist, dcount, pin, ev, cap, fallback_slots, mark)

    def _fallback(self, q, k, dlist, dcount, n_worst, fallback_slots, idx, dst):
        """Exact answers for the deferred queries: a block of F slots per rank (the same F everywhere), answered by
        every rank over its OWNED points, merged by (distance, global index)."""
        ops, G, ax, D = self.ops, self.world, self.axis, self.dim
        C, lib, check, st = ops.C, ops.lib, ops._abi.check, ops._st()
        dev = q.device
        F = max(int(fallback_slots), 1)
        while F < n_worst:
            F *= 2
        F = min(F, max(q.shape[0], 1))
        qf = torch.empty((F, D), dtype=torch.float64, device=dev)
        check(lib, lib.pomp_shard_gather_deferred(C.c_void_p(q.data_ptr()), D, C.c_void_p(dlist.data_ptr()),
                                                  C.c_void_p(dcount.data_ptr()), F,
                                                  C.c_void_p(self.pad_pts[self.rank].data_ptr()),
                                                  C.c_void_p(qf.data_ptr()), C.c_void_p(st)))
        allq = torch.empty((G * F, D), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allq, qf, group=self.group)
        ia, da = ops.local_knn(allq, k)
        reca, _ = ops.finalize(allq, ax, self.lo_cov, self.hi_cov, k, ia, da, self.l2g, self.halo_flag, True, False)
        got = torch.empty_like(reca)
        dist.all_to_all_single(got, reca, group=self.group)
        mi, md = ops.merge(got.view(G, F, k, 2), k)
        check(lib, lib.pomp_shard_scatter_deferred(C.c_void_p(mi.data_ptr()), C.c_void_p(md.data_ptr()), k,
                                                   C.c_void_p(dlist.data_ptr()), C.c_void_p(dcount.data_ptr()), F,
                                                   C.c_void_p(idx.data_ptr()), C.c_void_p(dst.data_ptr()),
                                                   C.c_void_p(st)))

    def neighbors(self, q, radius):
        """Collective radius query: CSR (offsets[nq_r + 1], global idx sorted ascending per query)."""
        G, ax = self.world, self.axis
        nq = q.shape[0]
        certified = radius <= self.halo
        ok = torch.tensor([1 if certified else 0], device=q.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()):
            qs, perm, counts = self.ops.route(q.contiguous(), ax, self.bounds, G)
            qr, rcounts = self._a2a(qs, counts)
            off, hits = self.ops.local_neighbors(qr, radius, self.inclusive_radius)
            ghits = self.l2g[hits]
            lens = off[1:] - off[:-1]
            # hits travel back grouped by source rank: per-rank hit totals, then the rows
            starts = [0]
            for c in rcounts:
                starts.append(starts[-1] + c)
            hit_counts = [int((off[starts[r + 1]] - off[starts[r]]).item()) for r in range(G)]
            lens_b = self._a2a_back(lens, rcounts, counts)
            hits_b, _ = self._a2a(ghits, hit_counts)
            # lens_b / hits_b are in routed (sorted-by-owner) order: rebuild the CSR in the caller's order
            off_s = torch.zeros(nq + 1, dtype=torch.int64, device=q.device)
            off_s[1:] = torch.cumsum(lens_b, 0)
            lens_o = torch.zeros(nq, dtype=torch.int64, device=q.device)
            lens_o[perm] = lens_b
            out_off = torch.zeros(nq + 1, dtype=torch.int64, device=q.device)
            out_off[1:] = torch.cumsum(lens_o, 0)
            seg = torch.repeat_interleave(torch.arange(nq, device=q.device), lens_b)
            within = torch.arange(hits_b.numel(), device=q.device) - off_s[seg]
            out = torch.empty_like(hits_b)
            out[out_off[perm[seg]] + within] = hits_b
            return out_off, out
        # radius beyond the halo: every rank answers every query over its OWNED points
        szs = torch.zeros(G, dtype=torch.int64, device=q.device)
        szs[self.rank] = nq
        dist.all_reduce(szs, group=self.group)
        m = int(szs.max().item())
        qf = torch.zeros((m, self.dim), dtype=torch.float64, device=q.device)
        qf[:nq] = q
        allq = torch.empty((G * m, self.dim), dtype=torch.float64, device=q.device)
        dist.all_gather_into_tensor(allq, qf, group=self.group)
        off, hits = self.ops.local_neighbors(allq, radius, self.inclusive_radius)
        keep = self.halo_flag[hits] == 0
        seg = torch.repeat_interleave(torch.arange(G * m, device=q.device), off[1:] - off[:-1])
        lens = torch.bincount(seg[keep], minlength=G * m)
        ghits = self.l2g[hits[keep]]
        lens_b = torch.empty_like(lens)
        dist.all_to_all_single(lens_b, lens, group=self.group)                # block g: rank g's counts for MY queries
        send_counts = [int(lens[r * m:(r + 1) * m].sum().item()) for r in range(G)]
        hits_b, _ = self._a2a(ghits, send_counts)
        L = lens_b.view(G, m)[:, :nq]
        tot = L.sum(0)
        out_off = torch.zeros(nq + 1, dtype=torch.int64, device=q.device)
        out_off[1:] = torch.cumsum(tot, 0)
        # hits_b: for g in ranks, for query j < m, that rank's hits; place them query-major, then sort each segment
        qid = torch.repeat_interleave(torch.arange(m, device=q.device).repeat(G), lens_b)
        sel = qid < nq
        key = qid[sel] * (self.n_global + 1) + hits_b[sel]
        order = torch.sort(key)[1]
        return out_off, hits_b[sel][order]

    # ------------------------------------------------------------------ mutation, routed to the owner
    def _owner(self, y):
        r = 0
        while r < self.world - 1 and not (y < self.bounds[r]):
            r += 1
        return r

    def holders(self, point):
        """Ranks that keep a copy of `point`: its owner and the neighbours whose halo it falls into."""
        y = float(point[self.axis])
        o = self._owner(y)
        out = [(o, 0)]
        if o > 0 and y - self.bounds[o - 1] <= self.halo:
            out.append((o - 1, 1))
        if o < self.world - 1 and self.bounds[o] - y <= self.halo:
            out.append((o + 1, 1))
        return out

    def add(self, point):
        """Same call on every rank (SPMD): appends the point under the next global index on its holders."""
        g = self.n_global
        self.n_global += 1
        for r, is_halo in self.holders(point):
            if r == self.rank:
                dev = self.local_pts.device
                self.local_pts = torch.cat([self.local_pts, torch.tensor([list(point)], dtype=torch.float64, device=dev)])
                self.l2g = torch.cat([self.l2g, torch.tensor([g], dtype=torch.int64, device=dev)])
                self.halo_flag = torch.cat([self.halo_flag, torch.tensor([is_halo], dtype=torch.uint8, device=dev)])
                self.ops.nn.add([list(point)]) if hasattr(self.ops, "nn") else self.ops.set_points(self.local_pts)
                if not is_halo:
                    self.n_owned += 1
        return g

    def local_index(self, gidx):
        """Local slot of a global insertion index on this rank, or -1."""
        if self.l2g.numel() == 0:
            return -1
        pos = int(torch.searchsorted(self.l2g, torch.tensor([gidx], dtype=torch.int64, device=self.l2g.device)).item())
        return pos if pos < self.l2g.numel() and int(self.l2g[pos].item()) == gidx else -1

    def remove(self, gidx):
        """Same call on every rank: tombstones every copy of the global index (owner + halo copies)."""
        li = self.local_index(gidx)
        if li >= 0:
            self.ops.nn.remove(li)
        return li >= 0

    def set_mask(self, global_mask):
        """global_mask: uint8 tensor over ALL global indices (1 = reject), or None."""
        if global_mask is None:
            self.ops.nn.set_mask(None)
        else:
            self.ops.nn.set_mask(global_mask[self.l2g].cpu().numpy())
