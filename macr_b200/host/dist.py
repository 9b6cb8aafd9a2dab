"""Multi-GPU plumbing of the one sub-path that shards with a real exchange step: full-catalogue
scoring.  One process per GPU (`torch.distributed`, NCCL over NVLink; gloo on CPU
for the host-logic tests).  The item table is row-partitioned into contiguous id ranges, query
rows are replicated, every rank emits its local top-K `[T,K] x (id, score)`, ONE all-gather
moves `T*K*8` bytes per rank and an on-device K-way merge (`macr_topk_merge`, the same kernel
that merges the in-GPU item chunks) produces the global list -- bit-identical to the unsharded
result because the order rule (score desc, lower id first) is a strict total order.

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
