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

The training step does not shard at benchmark sizes (a ~60 us step): ranks run replicas.
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
