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
is the whole exchange, everything else is the single-GPU step on the local slice; with one rank
per GPU the exchange is a fused push over NVLink peer memory instead (csrc/shard.cu).
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
    """Holds this rank's item shard (rows + gates) and scores query users against it.

    With one rank per GPU (NCCL group) the exchange of the shards' candidates is ONE kernel over
    NVLink peer memory (`macr_topk_merge_peers`): every rank writes its `[T,K]` candidates into a
    CUDA-IPC buffer, a flag barrier, then rank j merges the rows `[j*Tb, (j+1)*Tb)` reading all G
    shards' lists straight from the peers' buffers and stores the merged rows into every rank's
    result buffer, a second flag barrier -- all-to-all, K-way merge and all-gather fused, no
    staging, no NCCL on the data path.  `exchange="nccl"` (or a failed peer mapping) keeps
    `merge_shard_candidates` (all-to-all + merge + all-gather over NCCL; gloo in the CPU tests)."""

    def __init__(self, item_table, w, rank=None, world=None, group=None, exchange="auto", shard_of=None):
        from .. import ops  # CUDA library: only needed on the device path

        self.ops, self.group = ops, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        if shard_of is not None:  # share another scorer's shard (e.g. to time a different exchange)
            self.lo, self.hi, self.items, self.sig_i = shard_of.lo, shard_of.hi, shard_of.items, shard_of.sig_i
        else:
            n_items = item_table.shape[0]
            b = item_shard_bounds(n_items, self.world)
            self.lo, self.hi = int(b[self.rank]), int(b[self.rank + 1])
            self.items = item_table[self.lo:self.hi].contiguous()
            self.sig_i = ops.score_gates(self.items, w) if self.hi > self.lo else \
                torch.zeros(0, dtype=torch.float32, device=item_table.device)
        self._p2p = None  # (T, K) -> peer-mapped buffers
        self._epoch = 0
        self.exchange = "none" if self.world == 1 else "nccl"
        self._want_p2p = (exchange in ("auto", "p2p") and self.world > 1 and self.items.is_cuda
                          and dist.is_initialized() and dist.get_backend(group) == "nccl")
        if exchange == "p2p" and not self._want_p2p:
            raise ValueError("exchange='p2p' needs world > 1, CUDA and one rank per GPU (NCCL group)")

    def _prepared(self, c, K):
        """tensor-core operands of this rank's shard: items and gates are fixed for the scorer's
        lifetime, so they are prepared once per c"""
        if not (self.items.is_cuda and self.ops.uses_tc(self.items.shape[0], K)):
            return None
        if getattr(self, "_tc", None) is None or self._tc.c != float(c):
            self._tc = self.ops.TcItems(self.items, self.sig_i, c)
        return self._tc

    def _peer_buffers(self, T, K):
        """[cand ids | cand scores | result ids | result scores | flags] in ONE peer-mapped allocation,
        (re)built when the query shape changes; every rank maps every peer's copy."""
        import ctypes as C

        if self._p2p is not None and self._p2p["shape"] == (T, K):
            return self._p2p
        self._release_p2p()
        ops, dev, n = self.ops, self.items.device, T * K
        buf = ops.IpcBuffer(4 * n * 4 + 256, dev)
        every = [None] * self.world
        dist.all_gather_object(every, buf.handle, group=self.group)
        tabs = [(C.c_void_p * self.world)() for _ in range(5)]
        peers, ok = [], 1
        try:
            for r, h in enumerate(every):
                base = buf.ptr if r == self.rank else ops.IpcBuffer.open_peer(h)
                if r != self.rank:
                    peers.append(base)
                for k in range(5):
                    tabs[k][r] = base + 4 * n * k
        except ops.MacrError:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        flat = buf.as_f32((4 * n + 64,))
        st = {"shape": (T, K), "buf": buf, "peers": peers, "tabs": tabs, "ok": int(flag.item()) == 1,
              "cand": (flat[:n].view(torch.int32).view(T, K), flat[n:2 * n].view(T, K)),
              "res": (flat[2 * n:3 * n].view(torch.int32).view(T, K), flat[3 * n:4 * n].view(T, K)),
              "err": torch.zeros(1, dtype=torch.int32, device=dev)}
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)  # flags are zeroed and mapped everywhere before the first signal
        self._p2p = st
        self.exchange = "p2p" if st["ok"] else "nccl"
        return st

    def _release_p2p(self):
        if self._p2p is None:
            return
        torch.cuda.synchronize(self.items.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)  # no peer may still be reading / writing this rank's buffers
        for p in self._p2p["peers"]:
            self.ops.IpcBuffer.close_peer(p)
        self._p2p["buf"].free()
        self._p2p = None

    def close(self):
        self._release_p2p()

    def topk(self, Uq, sig_u, c, mask_rowptr, mask_col, K):
        ops, T = self.ops, Uq.shape[0]
        st = self._peer_buffers(T, K) if self._want_p2p and T > 0 else None
        if st is None or not st["ok"]:
            ids, sc = ops.score_topk(Uq, self.items, self.sig_i, sig_u, c, mask_rowptr, mask_col, K,
                                     item_id_offset=self.lo, prepared=self._prepared(c, K))
            if self.world == 1:
                return ids, sc
            return merge_shard_candidates(ops, ids, sc, self.world, self.group)
        ops.score_topk(Uq, self.items, self.sig_i, sig_u, c, mask_rowptr, mask_col, K, item_id_offset=self.lo,
                       out=st["cand"], prepared=self._prepared(c, K))
        t = st["tabs"]
        self._epoch += 1
        ops.shard_barrier(t[4], self.rank, self.world, self._epoch, st["err"])  # every shard's candidates are complete
        Tb = (T + self.world - 1) // self.world
        row0 = min(T, self.rank * Tb)
        ops.topk_merge_peers(t[0], t[1], t[2], t[3], self.world, K, row0, min(T, row0 + Tb) - row0)
        self._epoch += 1
        ops.shard_barrier(t[4], self.rank, self.world, self._epoch, st["err"])  # merged rows landed everywhere
        return st["res"][0].clone(), st["res"][1].clone()  # the buffers are reused by the next call

    def check_peers(self):
        if self._p2p is not None:
            e = int(self._p2p["err"].item())
            if e:
                raise self.ops.MacrError(f"rank {self.rank}: peer {e - 1} never reached a scoring barrier")


def merge_shard_candidates(ops, ids, scores, world, group=None):
    """Per-shard top-K lists `[T,K]` (ids, scores) -> the global `[T,K]` on every rank.

    Row blocks, not replicas: rank j merges rows `[j*Tb, (j+1)*Tb)` only.  One all-to-all hands it
    those rows of every shard's candidates (`T*K*8/G` bytes from each peer instead of the whole
    `T*K*8`), the on-device K-way merge (`macr_topk_merge`) runs on `Tb = ceil(T/G)` rows, and one
    all-gather of the merged blocks assembles the result -- `2*T*K*8` bytes received per rank
    instead of `G*T*K*8`, and 1/G of the merge work."""
    T, K = ids.shape
    Tb = (T + world - 1) // world
    both = torch.empty((world * Tb, 2 * K), dtype=torch.int32, device=ids.device)
    both[:T, :K] = ids
    both[:T, K:] = scores.view(torch.int32)
    if world * Tb > T:  # padding rows: empty lists
        both[T:, :K] = -1
        both[T:, K:] = torch.tensor(float("-inf"), dtype=torch.float32, device=ids.device).view(torch.int32)
    mine = torch.empty_like(both)  # [G, Tb, 2K]: block j = shard j's candidates for my rows
    dist.all_to_all_single(mine, both, group=group)
    mine = mine.view(world, Tb, 2 * K)
    mi, ms = ops.topk_merge(mine[:, :, :K].contiguous(), mine[:, :, K:].contiguous().view(torch.float32))
    out = torch.empty((world * Tb, 2 * K), dtype=torch.int32, device=ids.device)
    dist.all_gather_into_tensor(out, torch.cat([mi, ms.view(torch.int32)], dim=1), group=group)
    return out[:T, :K].contiguous(), out[:T, K:].contiguous().view(torch.float32)


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
        if getattr(self, "_tc", None) is None or self._tc.c != float(c):
            self._tc = self.ops.TcItems(self.items, self.sig_i, c) \
                if self.items.is_cuda and self.ops.uses_tc(self.items.shape[0], K) else None
        return self.ops.score_topk(Uq[lo:hi].contiguous(), self.items, self.sig_i,
                                   sig_u[lo:hi].contiguous(), c, mrp, mask_col, K,
                                   prepared=self._tc if hi > lo else None)

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

    One step = one exchange + the unchanged single-GPU step graph on the local slice.  Behind its
    owned rows every local table carries ghost rows (two parities, csrc/shard.cu); batch position
    b of the user / pos / neg column lives in ghost b / b / B+b.  Two transports for the exchange:

    * ``push`` (one rank per GPU, NVLink): INSIDE the captured step graph ONE kernel renumbers
      the ids (owned -> `id - lo`, foreign -> ghost slot) and stores every owned row straight
      into the ghost slot of every peer's table through CUDA-IPC mapped peer memory -- gather and
      all-gather fused, no staging buffer, no reduction.  The dense sweep of the owned rows starts
      at once; a one-warp flag barrier over peer memory sits on the gather branch only, so the
      HBM-bound part of the step never waits for a peer and an epoch is n graph replays.
    * ``allreduce``: `macr_shard_pack` writes the owned rows (zeros elsewhere) into `ex[3B,64]`,
      ONE all-reduce (sum; exactly one owner per row, so the sum is exact) hands every rank all 3B
      rows, `macr_shard_unpack` copies them into the ghost slots.  NCCL, or gloo for the tests.

    Then `macr_mf_trainer_step` runs as on one GPU.  Dots, the B x B grid, the losses and the
    gradients of `w` / `w_user` depend on batch positions only, so they come out identical on
    every rank (replicated work); row gradients and Adam touch owned rows exactly as the
    single-GPU step does (same segments, same summation order) and the dense sweep covers only
    the local slice -- the HBM traffic per rank is 1/G of the single-GPU step's.  Ghost rows
    receive meaningless updates and are overwritten before they are read again.

    The owned slices are bit-identical to the corresponding rows of a single-GPU run
    (tests/test_gpu_dist.py, tests/test_gpu_multi.py).  Batches must be identical on every rank."""

    def __init__(self, U, I, w, wu, hp, max_batch, rank=None, world=None, device="cuda:0", group=None,
                 ops_module=None, exchange="auto"):
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
        self.desc = ops.ShardDesc(self.rank, self.world, self.u_lo, self.u_hi, self.i_lo, self.i_hi, max_batch)
        dev = torch.device(device)
        if exchange not in ("auto", "push", "allreduce"):
            raise ValueError(f"exchange {exchange!r}")
        want_push = (exchange != "allreduce" and self.world > 1 and dev.type == "cuda"
                     and dist.is_initialized() and dist.get_backend(group) == "nccl")
        if exchange == "push" and not want_push:
            raise ValueError("exchange='push' needs world > 1, CUDA and one rank per GPU (NCCL group)")
        Uloc, self._ipc_u = ops.shard_table(self.n_lu + 2 * max_batch, dev, peer_mappable=want_push)
        Iloc, self._ipc_i = ops.shard_table(self.n_li + 4 * max_batch, dev, peer_mappable=want_push)
        as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, np.float32))
        Uloc[: self.n_lu].copy_(as_t(U[self.u_lo:self.u_hi]))
        Iloc[: self.n_li].copy_(as_t(I[self.i_lo:self.i_hi]))
        self.trainer = ops.MFTrainer(Uloc, Iloc, w, wu, hp, max_batch=max_batch, device=device)
        self.dev = self.trainer.dev
        self._local3 = torch.zeros(3 * max_batch, dtype=torch.int32, device=self.dev)
        self._ex = torch.zeros((3 * max_batch, Uloc.shape[1]), dtype=torch.float32, device=self.dev)
        self._step, self._peers = 0, []
        self._epoch_ids = self._epoch_losses = None
        self.exchange = "allreduce" if self.world > 1 else "none"
        if want_push:
            self._open_peers(strict=exchange == "push")

    # ---- peer memory (push transport) ---------------------------------------------------------
    def _open_peers(self, strict):
        import ctypes as C

        ops = self.ops
        mine = (self._ipc_u.handle, self._ipc_i.handle, self.trainer.ipc_export(), self.n_lu, self.n_li)
        every = [None] * self.world
        dist.all_gather_object(every, mine, group=self.group)
        pu, pi, pf = ((C.c_void_p * self.world)() for _ in range(3))
        ok = 1
        try:
            for r, (hu, hi, hf, n_lu, n_li) in enumerate(every):
                if r == self.rank:
                    continue
                bu, bi, bf = (ops.IpcBuffer.open_peer(h) for h in (hu, hi, hf))
                self._peers += [bu, bi, bf]
                pu[r], pi[r], pf[r] = bu + n_lu * 256, bi + n_li * 256, bf  # ghost bases: row n_local
        except ops.MacrError:
            if strict:
                raise
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)  # every rank mapped every peer?
        if int(flag.item()) == 1:
            self.trainer.shard(self.desc, pu, pi, pf)  # the exchange now runs inside the step graph
            self.exchange = "push"
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)  # flags are zeroed and mapped everywhere before the first push

    def check_peers(self):
        """Raise if a flag barrier timed out (a peer died); synchronises."""
        if self.exchange == "push":
            e = self.trainer.peer_error()
            if e:
                raise self.ops.MacrError(f"rank {self.rank}: peer {e - 1} never reached the exchange barrier")

    # ---- one step -----------------------------------------------------------------------------
    def exchange_rows(self, ids3, B):
        """all-reduce transport: ids3 = int32 device tensor [3B] = users | pos | neg, GLOBAL ids,
        identical on every rank.  Leaves the renumbered ids in `self._local3[:3B]` and every foreign
        row in its ghost slot.  (Push transport: the captured step does all of this itself.)"""
        ops, t, par = self.ops, self.trainer.tab, self._step & 1
        self._step += 1
        ex = self._ex[: 3 * B]
        ops.shard_pack(t.U, t.I, self.desc, ids3, B, par, self._local3, ex)
        if self.world > 1:  # every row has exactly one owner: x + 0 + ... is exact
            if ex.is_cuda and dist.get_backend(self.group) == "gloo":
                host = ex.cpu()  # gloo (several ranks sharing one GPU in the tests): stage through the host
                dist.all_reduce(host, group=self.group)
                ex.copy_(host)
            else:
                dist.all_reduce(ex, group=self.group)  # NCCL over NVLink, in the step's stream order
            ops.shard_unpack(t.U, t.I, self.desc, ex, B, par)

    def step_ids3(self, ids3, B, loss_out=None):
        """One training step on the [3B] global ids.  Returns the device tensor [loss, mf, reg,
        L_ori] (identical on every rank), written to `loss_out` ([1,4] device) if given; no sync."""
        if self.exchange == "push":  # renumbering, peer stores and barrier are part of the step graph
            if loss_out is not None:
                return self.trainer.run(ids3[: 3 * B].view(1, 3, B), loss_out)
            return self.trainer.step_device(ids3[:B], ids3[B:2 * B], ids3[2 * B:3 * B])
        self.exchange_rows(ids3, B)
        l3 = self._local3
        if loss_out is not None:
            return self.trainer.run(l3[: 3 * B].view(1, 3, B), loss_out)
        return self.trainer.step_device(l3[:B], l3[B:2 * B], l3[2 * B:3 * B])

    def step_device(self, users, pos, neg):
        """users / pos / neg: int32 device tensors [B] of GLOBAL ids, identical on every rank."""
        return self.step_ids3(torch.cat([users, pos, neg]), users.numel())

    def run(self, batches, losses=None):
        """Epoch mode: int32 device [n,3,B] GLOBAL ids -> device losses [n,4].  Push transport: n
        replays of one captured graph, no host work between the steps."""
        n, three, B = batches.shape
        if losses is None:
            losses = torch.empty((n, 4), dtype=torch.float32, device=self.dev)
        if self.exchange == "push":
            return self.trainer.run(batches, losses)
        for s in range(n):
            self.step_ids3(batches[s].reshape(-1), B, losses[s:s + 1])
        return losses

    def run_host(self, batches_host, losses_host=None):
        """Epoch call with HOST buffers: `batches_host` int32 [n,3,B] (pinned recommended, identical
        on every rank) -> float32 [n,4] host losses; one H2D, n exchanges + steps, one D2H, one sync."""
        n, three, B = batches_host.shape
        assert three == 3 and batches_host.dtype == torch.int32 and not batches_host.is_cuda
        if self.exchange == "push":  # macr_mf_trainer_run_host: the whole epoch behind one C call
            return self.trainer.run_host(batches_host, losses_host)
        if self._epoch_ids is None or self._epoch_ids.shape[0] < n or self._epoch_ids.shape[2] != B:
            self._epoch_ids = torch.empty((n, 3, B), dtype=torch.int32, device=self.dev)
            self._epoch_losses = torch.empty((n, 4), dtype=torch.float32, device=self.dev)
        ids, losses = self._epoch_ids[:n], self._epoch_losses[:n]
        ids.copy_(batches_host, non_blocking=True)
        for s in range(n):
            self.step_ids3(ids[s].view(-1), B, losses[s:s + 1])
        if losses_host is None:
            losses_host = torch.empty((n, 4), dtype=torch.float32)
        losses_host.copy_(losses, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return losses_host

    def local_tables(self):
        """views of the owned rows: (U[u_lo:u_hi], I[i_lo:i_hi]) and their Adam slots by name"""
        t = self.trainer.tab
        return {"U": t.U[: self.n_lu], "mU": t.mU[: self.n_lu], "vU": t.vU[: self.n_lu],
                "I": t.I[: self.n_li], "mI": t.mI[: self.n_li], "vI": t.vI[: self.n_li],
                "w": t.w, "wu": t.wu}

    def close(self):
        if self.trainer is None:
            return
        if self.world > 1 and self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
            if dist.is_initialized():
                dist.barrier(group=self.group)  # no peer may still be pushing into this rank's tables
        self.trainer.close()
        self.trainer = None
        for p in self._peers:
            self.ops.IpcBuffer.close_peer(p)
        self._peers = []
        for buf in (self._ipc_u, self._ipc_i):
            if buf is not None:
                buf.free()


def balanced_bounds(weights, world):
    """Contiguous ranges of ~equal total weight: rank r owns [b[r], b[r+1]).  A Zipf catalogue sorted
    by id puts most nonzeros into the first item rows; equal row counts would leave rank 0 with
    half of the SpMM work."""
    c = np.concatenate([[0], np.cumsum(np.asarray(weights, np.int64))])
    b = np.searchsorted(c, c[-1] * np.arange(world + 1) / world, side="left").astype(np.int64)
    b[0], b[-1] = 0, len(weights)
    return np.maximum.accumulate(b)


def partition_adjacency(rowptr, col, val, n_users, u_lo, u_hi, i_lo, i_hi):
    """1-D row partition of the normalised adjacency: the CSR of the rows a rank
    owns -- user rows [u_lo,u_hi) and item rows n_users + [i_lo,i_hi) -- with the full N+1 row
    pointer (rows of other ranks are empty) and GLOBAL column ids."""
    rowptr = np.asarray(rowptr, np.int64)
    n = len(rowptr) - 1
    own = np.zeros(n, bool)
    own[u_lo:u_hi] = True
    own[n_users + i_lo:n_users + i_hi] = True
    deg = np.diff(rowptr)
    new_rp = np.zeros(n + 1, np.int64)
    new_rp[1:] = np.cumsum(np.where(own, deg, 0))
    sel = np.repeat(own, deg)
    return (new_rp.astype(np.int32), np.ascontiguousarray(np.asarray(col)[sel], np.int32),
            np.ascontiguousarray(np.asarray(val)[sel], np.float32))


class RowShardedLGCNTrainer:
    """LightGCN `bceboth` training with the adjacency, the propagation and the dense Adam
    row-partitioned over the ranks (SURVEY 8e row "LightGCN SpMM"; LightGCN.py:257-269,297-305).

    Rank r owns contiguous user and item id ranges: it holds the nonzeros of those rows only,
    computes every propagation layer (forward and backward) for them and applies Adam to them.
    Each [N,64] layer buffer is full-size on every rank; after a layer the owned rows are stored
    straight into the peers' buffers over NVLink (CUDA-IPC mapped peer memory, a flag barrier after
    the stores) -- the all-gather of E_k, inside the step's CUDA graph, no NCCL on the data path.
    Dots, the B x B grid and the batch's row gradients are replicated (batch positions only).
    Owned rows, losses, w and w_user are bit-identical to the single-GPU trainer: a row's segment
    plan depends on the row alone (tests/test_gpu_multi.py).  One rank per GPU (NCCL group for the
    handle exchange); batches must be identical on every rank."""

    def __init__(self, rowptr, col, val, U, I, w, wu, n_layers, hp, max_batch, rank=None, world=None,
                 device="cuda:0", group=None):
        import ctypes as C

        from .. import ops

        self.ops, self.group = ops, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.n_users, self.n_items = U.shape[0], I.shape[0]
        # owned ranges balanced by work: a row costs its nonzeros (256-byte gathers) + ~8 (Adam, epilogue)
        deg = np.diff(np.asarray(rowptr, np.int64))
        ub = balanced_bounds(deg[: self.n_users] + 8, self.world)
        ib = balanced_bounds(deg[self.n_users:] + 8, self.world)
        self.u_lo, self.u_hi = int(ub[self.rank]), int(ub[self.rank + 1])
        self.i_lo, self.i_hi = int(ib[self.rank]), int(ib[self.rank + 1])
        self.dev = torch.device(device)
        rp, cl, vl = partition_adjacency(rowptr, col, val, self.n_users, self.u_lo, self.u_hi, self.i_lo, self.i_hi)
        self.local_nnz = int(len(cl))
        d = U.shape[1]
        self._ipc_u = ops.IpcBuffer(self.n_users * d * 4, self.dev)
        self._ipc_i = ops.IpcBuffer(self.n_items * d * 4, self.dev)
        tU, tI = self._ipc_u.as_f32((self.n_users, d)), self._ipc_i.as_f32((self.n_items, d))
        as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, np.float32))
        tU.copy_(as_t(U))  # every rank starts from the full tables; from then on owners refresh their rows
        tI.copy_(as_t(I))
        self.trainer = ops.LGCNTrainer(rp, cl, vl, tU, tI, w, wu, n_layers, hp, max_batch, device=self.dev)
        self._peers = []
        desc = ops.ShardDesc(self.rank, self.world, self.u_lo, self.u_hi, self.i_lo, self.i_hi, max_batch)
        pu, pi, pe, pt, pf = ((C.c_void_p * self.world)() for _ in range(5))
        if self.world > 1:
            mine = [self._ipc_u.handle, self._ipc_i.handle] + self.trainer.ipc_export()
            every = [None] * self.world
            dist.all_gather_object(every, mine, group=self.group)
            ok = 1
            try:
                for r, hs in enumerate(every):
                    if r == self.rank:
                        continue
                    ptrs = [ops.IpcBuffer.open_peer(h) for h in hs]
                    self._peers += ptrs
                    pu[r], pi[r], pe[r], pt[r], pf[r] = ptrs
            except ops.MacrError:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) != 1:  # fail on EVERY rank, so nobody is left waiting in a barrier
                self.close()
                raise ops.MacrError("row-partitioned LightGCN needs CUDA-IPC peer memory between the ranks "
                                    "(cudaIpcOpenMemHandle failed on at least one rank)")
        self.trainer.shard(desc, pu, pi, pe, pt, pf)
        if self.world > 1:
            torch.cuda.synchronize(self.dev)
            dist.barrier(group=self.group)  # everybody is mapped before the first peer store

    def step_device(self, users, pos, neg, train=True):
        return self.trainer.step_device(users, pos, neg, train)

    def run(self, batches, train=True, losses=None):
        return self.trainer.run(batches, train, losses)

    def run_host(self, batches_host, losses_host=None):
        return self.trainer.run_host(batches_host, losses_host)

    def embeddings(self):
        """(users [U,64], items [I,64]) propagated tables, complete on every rank."""
        return self.trainer.embeddings()

    def check_peers(self):
        e = self.trainer.peer_error()
        if e:
            raise self.ops.MacrError(f"rank {self.rank}: peer {e - 1} never reached an exchange barrier")

    def local_tables(self):
        t = self.trainer.tab
        return {"U": t.U[self.u_lo:self.u_hi], "mU": t.mU[self.u_lo:self.u_hi], "vU": t.vU[self.u_lo:self.u_hi],
                "I": t.I[self.i_lo:self.i_hi], "mI": t.mI[self.i_lo:self.i_hi], "vI": t.vI[self.i_lo:self.i_hi],
                "w": t.w, "wu": t.wu}

    def close(self):
        if self.trainer is None:
            return
        torch.cuda.synchronize(self.dev)
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)  # no peer may still be storing into this rank's buffers
        self.trainer.close()
        self.trainer = None
        for p in self._peers:
            self.ops.IpcBuffer.close_peer(p)
        self._peers = []
        self._ipc_u.free()
        self._ipc_i.free()
