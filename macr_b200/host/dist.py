"""Multi-GPU plumbing of the one sub-path that shards with a real exchange step: full-catalogue
scoring.  One process per GPU (`torch.distributed`, NCCL over NVLink; gloo on CPU
for the host-logic tests).  The item table is row-partitioned into contiguous id ranges, query
rows are replicated, every rank emits its local top-K `[T,K] x (id, score)`, ONE all-gather
moves `T*K*8` bytes per rank and an on-device K-way merge (`macr_topk_merge`, the same kernel
that merges the in-GPU item chunks) produces the global list -- bit-identical to the unsharded
result because the order rule (score desc, lower id first) is a strict total order.

Two partitionings:

* ``ShardedScorer`` -- ITEMS partitioned: the layout for a catalogue that does not fit (or is not
  wanted) on every GPU.  One exchange step: the all-gather of the per-shard candidates.
* ``UserShardedScorer`` -- QUERY USERS partitioned, item table replicated: no exchange on the data
  path at all (the catalogues of the reference are 0.2 .. 10 MB), every rank scores its slice of
  the query users against the full catalogue; one all-gather assembles the `[T,K]` result.  The
  per-row stages of the tensor-core pipeline (threshold, re-rank) shrink with the slice, so this
  is the layout that scales; `bench.py` reports both.

The training step does not shard profitably at benchmark sizes (a ~60 us step): `bench.py` runs
replicas.  For tables whose dense-Adam traffic dominates (SURVEY 8e rows "dense Adam sweep" and
"gather + grid + row grads"; config 5: 16.9 GB per step on one GPU) ``RowShardedMFTrainer``
row-partitions both tables with their Adam slots: one all-reduce of the 3B gathered rows per step
is the whole exchange, everything else is the single-GPU step on the local slice.
"""
import numpy as np
import torch
import torch.distributed as dist


def item_shard_bounds(n_items, world):
    """Contiguous, balanced item-id ranges: rank r owns [b[r], b[r+1])."""
    return np.linspace(0, n_items, world + 1).astype(np.int64)


def all_gather_candidates(ids, scores, group=None):
    """[T,K] per rank -> ([G,T,K] ids, [G,T,K] scores) on every rank, shard order = rank order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids.unsqueeze(0), scores.unsqueeze(0)
    T, K = ids.shape
    gi = torch.empty((world * T, K), dtype=ids.dtype, device=ids.device)  # rank-major concat
    gs = torch.empty((world * T, K), dtype=scores.dtype, device=scores.device)
    dist.all_gather_into_tensor(gi, ids.contiguous(), group=group)
    dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
    return gi.view(world, T, K), gs.view(world, T, K)


class ShardedScorer:
    """Holds this rank's item shard (rows + gates) and scores query users against it."""

    def __init__(self, item_table, w, rank=None, world=None, group=None):
        from .. import ops  # CUDA library: only needed on the device path

        self.ops, self.group = ops, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        n_items = item_table.shape[0]
        b = item_shard_bounds(n_items, self.world)
        self.lo, self.hi = int(b[self.rank]), int(b[self.rank + 1])
        self.items = item_table[self.lo:self.hi].contiguous()
        self.sig_i = ops.score_gates(self.items, w) if self.hi > self.lo else \
            torch.zeros(0, dtype=torch.float32, device=item_table.device)

    def topk(self, Uq, sig_u, c, mask_rowptr, mask_col, K):
        ids, sc = self.ops.score_topk(Uq, self.items, self.sig_i, sig_u, c, mask_rowptr, mask_col, K,
                                      item_id_offset=self.lo)
        if self.world == 1:
            return ids, sc
        gi, gs = all_gather_candidates(ids, sc, self.group)
        return self.ops.topk_merge(gi, gs)


def user_shard_bounds(n_rows, world):
    """Contiguous, balanced query-row ranges: rank r owns [b[r], b[r+1])."""
    return np.linspace(0, n_rows, world + 1).astype(np.int64)


def all_gather_rows(local, n_rows, world, rank, group=None):
    """Ragged row-slices `[b[r]:b[r+1], ...]` per rank -> the full `[n_rows, ...]` tensor on every
    rank (slices padded to the longest one for the collective)."""
    if world == 1:
        return local
    b = user_shard_bounds(n_rows, world)
    longest = int(np.max(np.diff(b)))
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    out = out.view((world, longest) + tuple(local.shape[1:]))
    return torch.cat([out[r, :int(b[r + 1] - b[r])] for r in range(world)], dim=0)


class UserShardedScorer:
    """Item table + gates replicated; rank r scores query rows [b[r], b[r+1]) of every call."""

    def __init__(self, item_table, w, rank=None, world=None, group=None):
        from .. import ops

        self.ops, self.group = ops, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.items = item_table.contiguous()
        self.sig_i = ops.score_gates(self.items, w)

    def local_rows(self, T):
        b = user_shard_bounds(T, self.world)
        return int(b[self.rank]), int(b[self.rank + 1])

    def topk_local(self, Uq, sig_u, c, mask_rowptr, mask_col, K):
        """Top-K of this rank's slice of the query rows: ([n_local,K] ids, scores)."""
        lo, hi = self.local_rows(Uq.shape[0])
        mrp = None
        if mask_rowptr is not None:
            mrp = mask_rowptr[lo:hi + 1].contiguous()  # absolute offsets into mask_col stay valid
        return self.ops.score_topk(Uq[lo:hi].contiguous(), self.items, self.sig_i,
                                   sig_u[lo:hi].contiguous(), c, mrp, mask_col, K)

    def topk(self, Uq, sig_u, c, mask_rowptr, mask_col, K):
        ids, sc = self.topk_local(Uq, sig_u, c, mask_rowptr, mask_col, K)
        if self.world == 1:
            return ids, sc
        # one collective: ids and score bit patterns side by side in an int32 [n, 2K] block
        both = torch.cat([ids, sc.view(torch.int32)], dim=1)
        full = all_gather_rows(both, Uq.shape[0], self.world, self.rank, self.group)
        return full[:, :K].contiguous(), full[:, K:].contiguous().view(torch.float32)


class RowShardedMFTrainer:
    """MF training with the two embedding tables AND their Adam slots row-partitioned over the
    ranks (contiguous id ranges, `user_shard_bounds` / `item_shard_bounds`).

    One step = one exchange + the unchanged single-GPU step graph on the local slice:

    1. every rank gathers the rows it owns among the batch's 3B ids (zeros elsewhere) and ONE
       all-reduce (sum; exactly one owner per row, so the sum is exact) hands every rank all 3B
       rows (`3*B*64*4` bytes: 3 MB at B = 4096);
    2. rows of other ranks are parked in ghost rows appended to the local tables
       (`n_local + position`), ids are renumbered (owned -> `id - lo`, foreign -> ghost slot);
    3. `macr_mf_trainer_step` runs as on one GPU.  Dots, the B x B grid, the losses and the
       gradients of `w` / `w_user` depend on batch positions only, so they come out identical on
       every rank (replicated work, ~25 us); row gradients and Adam touch owned rows exactly as the
       single-GPU step does (same segments, same summation order) and the dense sweep covers only
       the local slice -- the HBM traffic per rank is 1/G of the single-GPU step's.  Ghost rows
       receive meaningless updates and are overwritten before they are read again.

    The owned slices are bit-identical to the corresponding rows of a single-GPU run
    (tests/test_gpu_dist.py).  Batches must be identical on every rank."""

    def __init__(self, U, I, w, wu, hp, max_batch, rank=None, world=None, device="cuda:0", group=None,
                 ops_module=None):
        if ops_module is None:  # the CUDA library; the CPU plumbing test injects a stand-in
            from .. import ops as ops_module
        ops = ops_module
        self.ops, self.group = ops, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.n_users, self.n_items, self.max_batch = U.shape[0], I.shape[0], max_batch
        ub, ib = user_shard_bounds(self.n_users, self.world), item_shard_bounds(self.n_items, self.world)
        self.u_lo, self.u_hi = int(ub[self.rank]), int(ub[self.rank + 1])
        self.i_lo, self.i_hi = int(ib[self.rank]), int(ib[self.rank + 1])
        self.n_lu, self.n_li = self.u_hi - self.u_lo, self.i_hi - self.i_lo
        d = U.shape[1]
        ghost = lambda rows: np.zeros((rows, d), np.float32)
        Uloc = np.concatenate([np.asarray(U[self.u_lo:self.u_hi], np.float32), ghost(max_batch)])
        Iloc = np.concatenate([np.asarray(I[self.i_lo:self.i_hi], np.float32), ghost(2 * max_batch)])
        self.trainer = ops.MFTrainer(Uloc, Iloc, w, wu, hp, max_batch=max_batch, device=device)
        self.dev = self.trainer.dev
        self._pos = torch.arange(2 * max_batch, dtype=torch.int32, device=self.dev)

    def _localize(self, ids, lo, hi, n_local, slot0):
        """global ids -> (local ids with foreign rows sent to ghost slots, ownership mask)"""
        own = (ids >= lo) & (ids < hi)
        ghost = n_local + slot0 + self._pos[: ids.numel()]
        return torch.where(own, ids - lo, ghost).to(torch.int32), own

    def step_device(self, users, pos, neg):
        """users / pos / neg: int32 device tensors [B] of GLOBAL ids, identical on every rank.
        Returns the device tensor [loss, mf, reg, L_ori] (identical on every rank); no sync."""
        ops, t = self.ops, self.trainer.tab
        B = users.numel()
        lu, own_u = self._localize(users, self.u_lo, self.u_hi, self.n_lu, 0)
        lp, own_p = self._localize(pos, self.i_lo, self.i_hi, self.n_li, 0)
        ln, own_n = self._localize(neg, self.i_lo, self.i_hi, self.n_li, B)
        rows = torch.cat([ops.gather_rows(t.U, lu), ops.gather_rows(t.I, torch.cat([lp, ln]))])
        own = torch.cat([own_u, own_p, own_n]).unsqueeze(1)
        ex = torch.where(own, rows, torch.zeros((), dtype=torch.float32, device=self.dev))
        if self.world > 1:  # every row has exactly one owner: x + 0 + ... is exact
            if ex.is_cuda and dist.get_backend(self.group) == "gloo":
                host = ex.cpu()  # gloo (several ranks sharing one GPU in the tests): stage through the host
                dist.all_reduce(host, group=self.group)
                ex.copy_(host)
            else:
                dist.all_reduce(ex, group=self.group)  # NCCL over NVLink, on the step's stream order
        t.U[self.n_lu:self.n_lu + B].copy_(ex[:B])           # ghost slots (owned positions unused)
        t.I[self.n_li:self.n_li + 2 * B].copy_(ex[B:])
        return self.trainer.step_device(lu, lp, ln)

    def local_tables(self):
        """views of the owned rows: (U[u_lo:u_hi], I[i_lo:i_hi]) and their Adam slots by name"""
        t = self.trainer.tab
        return {"U": t.U[: self.n_lu], "mU": t.mU[: self.n_lu], "vU": t.vU[: self.n_lu],
                "I": t.I[: self.n_li], "mI": t.mI[: self.n_li], "vI": t.vI[: self.n_li],
                "w": t.w, "wu": t.wu}

    def close(self):
        self.trainer.close()
