#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs[2..4] (bench.py measures configs[1]).

  python dev/bench_configs.py [--what lgcn,ml10m,synth] [--synth-users N]
  torchrun --nproc-per-node G dev/bench_configs.py --what ml10m,synth      (scoring, users/G)

Prints one JSON line per measurement (rank 0).  Device times are CUDA events on the launching
stream, max over ranks; every input is synthetic with the shapes of SURVEY.md 8(d).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

D = 64


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def xavier(rng, rows, scale=1.0):
    lim = np.sqrt(6.0 / (rows + D))
    return (rng.uniform(-lim, lim, (rows, D)) * scale).astype(np.float32)


def synth_graph(n_users, n_items, n_edges, seed):
    """random bipartite train graph with Zipf-popular items -> D^-1/2 A D^-1/2 CSR (float32)"""
    import scipy.sparse as sp

    rng = np.random.RandomState(seed)
    u = rng.randint(0, n_users, int(n_edges * 1.15))
    ranks = np.arange(1, n_items + 1, dtype=np.float64)
    cdf = np.cumsum(1.0 / ranks)
    cdf /= cdf[-1]
    i = np.minimum(np.searchsorted(cdf, rng.rand(len(u))), n_items - 1)
    key = np.unique(u.astype(np.int64) * n_items + i)[:n_edges]
    u, i = key // n_items, key % n_items
    R = sp.csr_matrix((np.ones(len(u), np.float32), (u, i)), shape=(n_users, n_items))
    A = sp.bmat([[None, R], [R.T, None]], format="csr", dtype=np.float32)
    deg = np.asarray(A.sum(1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, -0.5)
    dinv[np.isinf(dinv)] = 0
    N = sp.diags(dinv).dot(A).dot(sp.diags(dinv)).tocsr().astype(np.float32)
    N.sort_indices()
    return N.indptr.astype(np.int32), N.indices.astype(np.int32), N.data.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="lgcn,ml10m,synth")
    ap.add_argument("--synth-users", type=int, default=262144)
    ap.add_argument("--synth-items", type=int, default=1000000)
    ap.add_argument("--lgcn-datasets", default="yelp2018,gowalla,ml_10m")
    ap.add_argument("--train-users", type=int, default=10_000_000)
    ap.add_argument("--train-items", type=int, default=1_000_000)
    args = ap.parse_args()
    what = set(args.what.split(","))

    import torch
    import torch.distributed as dist

    from macr_b200 import ops
    from macr_b200.host.dist import UserShardedScorer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk, pk_kind = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def tmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return tmax(e0.elapsed_time(e1) * 1e-3) / reps

    def emit(**kw):
        if rank == 0:
            kw.update(n_gpus=world, data="synthetic", peak_kind=pk_kind)
            print(json.dumps(kw), flush=True)

    # ---------------- config 3: LightGCN yelp2018 shapes, L = 2, one GPU ----------------
    lgcn_sets = [d for d in args.lgcn_datasets.split(",") if d] if ("lgcn" in what and world == 1) else []
    for ds in lgcn_sets:
        B, L = 4096, 2
        path = os.path.join(ROOT, "data", ds)
        if os.path.exists(os.path.join(path, "train.txt")):
            import contextlib
            import io

            from macr_b200.host.data_lgcn import Data

            with contextlib.redirect_stdout(io.StringIO()):
                data = Data(path, B)
                rowptr, col, val = data.adj_csr("pre")  # the reference's own D^-1/2 A D^-1/2
            U_n, I_n, src = data.n_users, data.n_items, "real adjacency (data/%s)" % ds
        else:
            U_n, I_n, n_train = 31668, 38048, 1371166
            rowptr, col, val = synth_graph(U_n, I_n, n_train, 1)
            src = "synthetic Zipf graph with yelp2018 shapes"
        N, nnz = U_n + I_n, len(col)
        max_deg = int(np.max(np.diff(rowptr)))
        rng = np.random.RandomState(2)
        Ue, Ie = xavier(rng, U_n), xavier(rng, I_n)
        w = rng.uniform(-0.3, 0.3, D).astype(np.float32)
        wu = rng.uniform(-0.3, 0.3, D).astype(np.float32)
        d_rp, d_col, d_val = (torch.from_numpy(x).to(dev) for x in (rowptr, col, val))
        X = torch.from_numpy(np.concatenate([Ue, Ie])).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def flushed(fn, reps=20):
            ts = []
            for _ in range(3):
                fn()
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3)
            return float(np.mean(ts))

        plan = ops.SpmmPlan(d_rp)  # static segment decomposition, built once per graph
        t_spmm = flushed(lambda: ops.spmm_csr(d_rp, d_col, d_val, X, plan=plan))
        t_spmm_stateless = flushed(lambda: ops.spmm_csr(d_rp, d_col, d_val, X), reps=5)
        spmm_bytes = 8.0 * nnz + 4.0 * (N + 1) + 8.0 * N * D
        gather_bytes = 8.0 * nnz + 4.0 * D * nnz + 4.0 * N * D  # every nonzero reads a 256-byte row (L2)
        dU, dI = torch.from_numpy(Ue).to(dev), torch.from_numpy(Ie).to(dev)
        t_prop = flushed(lambda: ops.lgcn_propagate(d_rp, d_col, d_val, dU, dI, L, plan=plan))
        hp = ops.HParams.make(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=B)
        tr = ops.LGCNTrainer(rowptr, col, val, Ue, Ie, w, wu, L, hp, max_batch=B, device=dev)
        nb = 64
        batches = np.empty((nb, 3, B), np.int32)
        for s in range(nb):
            batches[s, 0] = rng.permutation(U_n)[:B]
            batches[s, 1] = rng.randint(0, I_n, B)
            batches[s, 2] = rng.randint(0, I_n, B)
        d_b = torch.from_numpy(batches).to(dev)
        losses = torch.zeros((nb, 4), dtype=torch.float32, device=dev)
        k = [0]

        def step():
            s = k[0] % nb
            tr.run(d_b[s:s + 1], True, losses[s:s + 1])
            k[0] += 1

        t_step = flushed(step, reps=40)
        step_bytes = 2 * L * spmm_bytes + 24.0 * D * N + 12.0 * D * B
        emit(config="MACR-LightGCN %s U=%d I=%d nnz(A)=%d max row %d L=2 d=64 B=4096 bceboth" %
                    (ds, U_n, I_n, nnz, max_deg), graph=src,
             l2="256 MiB memset before every timed call",
             spmm={"ms": 1e3 * t_spmm, "ms_stateless_kernel": 1e3 * t_spmm_stateless,
                   "bytes": spmm_bytes, "achieved_gbs": spmm_bytes / t_spmm / 1e9,
                   "peak_gbs": pk["hbm_gbs"], "frac": spmm_bytes / t_spmm / 1e9 / pk["hbm_gbs"],
                   "l2_gather_bytes": gather_bytes, "l2_gather_gbs": gather_bytes / t_spmm / 1e9,
                   "note": "algorithmic bytes assume X is read once; each nonzero actually gathers a "
                           "256-byte row from L2 (X is L2-resident), which is what bounds the kernel"},
             propagate={"ms": 1e3 * t_prop, "layers": L},
             step={"ms": 1e3 * t_step, "interactions_per_s": B / t_step, "launches": tr.launches_per_step,
                   "bytes": step_bytes, "achieved_gbs": step_bytes / t_step / 1e9,
                   "frac": step_bytes / t_step / 1e9 / pk["hbm_gbs"]},
             final_loss=float(losses[(k[0] - 1) % nb, 0].item()))
        tr.close()

    # ---------------- scoring shapes: config 4 (ml_10m) and config 5 (synthetic) ----------------
    def scoring(name, T_q, n_items, avg_mask, reps):
        rng = np.random.RandomState(7)
        It = torch.from_numpy(xavier(rng, n_items, 10.0)).to(dev)
        w = torch.from_numpy(rng.uniform(-0.3, 0.3, D).astype(np.float32)).to(dev)
        wu = torch.from_numpy(rng.uniform(-0.3, 0.3, D).astype(np.float32)).to(dev)
        sc = UserShardedScorer(It, w, rank=rank, world=world)
        lo, hi = sc.local_rows(T_q)
        n_loc = hi - lo
        # this rank's slice only (the full query set of config 5 does not need to exist anywhere)
        Uq = torch.from_numpy(xavier(np.random.RandomState(100 + rank), n_loc, 10.0)).to(dev)
        su = ops.score_gates(Uq, wu)
        cnt = np.maximum(1, np.random.RandomState(200 + rank).poisson(avg_mask, n_loc)).astype(np.int64)
        rp = np.zeros(n_loc + 1, np.int64)
        rp[1:] = np.cumsum(cnt)
        colm = np.random.RandomState(300 + rank).randint(0, n_items, int(rp[-1])).astype(np.int32)
        # sorted within each row (duplicates are harmless for a mask)
        order = np.lexsort((colm, np.repeat(np.arange(n_loc), cnt)))
        colm = colm[order]
        d_rp, d_col = torch.from_numpy(rp.astype(np.int32)).to(dev), torch.from_numpy(colm).to(dev)
        stats = torch.zeros(2, dtype=torch.int64, device=dev)

        def once():
            return ops.score_topk(Uq, sc.items, sc.sig_i, su, 40.0, d_rp, d_col, 20)

        t = timed(once, reps, warm=2)
        ops.score_topk_tc(Uq, sc.items, sc.sig_i, su, 40.0, d_rp, d_col, 20, stats=stats)
        st = stats.cpu().tolist()
        # spot parity: a few rows against the exact fp32 kernel
        ids, scs = once()
        sel = torch.arange(0, n_loc, max(1, n_loc // 256), device=dev)[:256]
        sub_rp = torch.zeros(len(sel) + 1, dtype=torch.int32, device=dev)
        lens = (d_rp[sel.long() + 1] - d_rp[sel.long()])
        sub_rp[1:] = torch.cumsum(lens, 0)
        sub_col = torch.cat([d_col[int(d_rp[r]):int(d_rp[r + 1])] for r in sel.tolist()])
        ei, es = ops.score_topk_exact(Uq[sel.long()].contiguous(), sc.items, sc.sig_i, su[sel.long()].contiguous(),
                                      40.0, sub_rp, sub_col, 20)
        same = bool((ei == ids[sel.long()]).all().item() and (es == scs[sel.long()]).all().item())
        n_pad = (n_items + 255) // 256 * 256
        flops = 2 * 2.0 * (D + 16) * T_q * n_pad
        emit(config=name, metric="full_catalog_scores_per_sec", value=T_q * n_items / t, unit="scores/s",
             ms_per_eval=1e3 * t, test_users=T_q, items=n_items, topk=20, sharding=f"users/{world}",
             rows_redone_by_exact_kernel_rank0=st[0], candidates_per_row_rank0=st[1] / max(1, n_loc - st[0]),
             spot_parity_vs_exact_fp32_kernel=same,
             roofline={"bound": "tensor", "achieved": flops / t / 1e12, "unit": "TFLOP/s",
                       "peak": pk["bf16_tflops"] * world, "frac": flops / t / 1e12 / (pk["bf16_tflops"] * world)})

    # ---------------- config 5, training side: MF step on 10 M x 1 M tables (HBM-bound) ----------------
    train_cfgs = []
    if "synth_train" in what and world == 1:
        train_cfgs.append(("synthetic (config 5 tables)", args.train_users, args.train_items, 8192))
    if "ml10m_train" in what and world == 1:
        train_cfgs.append(("ml_10m-shape (config 4 tables, L2-resident: back-to-back replay)", 69166, 8790, 8192))
    for tname, U_n, I_n, B in train_cfgs:
        gen = torch.Generator(device=dev).manual_seed(1)
        lim_u, lim_i = float(np.sqrt(6.0 / (U_n + D))), float(np.sqrt(6.0 / (I_n + D)))
        # tables are created on the device (8.45 GB of var/m/v); the trainer adopts host arrays, so
        # it is built on 1-row placeholders and its table tensors are swapped before the first step
        rng = np.random.RandomState(3)
        w = rng.uniform(-0.3, 0.3, D).astype(np.float32)
        wu = rng.uniform(-0.3, 0.3, D).astype(np.float32)
        hp = ops.HParams.make(lr=1e-3, alpha=1e-3, beta=1e-3, decay=1e-5, batch_size=B)
        Uh = np.zeros((U_n, D), np.float32)
        Ih = np.zeros((I_n, D), np.float32)
        tr = ops.MFTrainer(Uh, Ih, w, wu, hp, max_batch=B, device=dev)
        del Uh, Ih
        t = tr.tab
        t.U.uniform_(-lim_u, lim_u, generator=gen)
        t.I.uniform_(-lim_i, lim_i, generator=gen)
        for m_, v_ in ((t.mU, t.vU), (t.mI, t.vI)):  # steady state: every row has non-zero moments
            m_.normal_(0.0, 1e-4, generator=gen)
            v_.uniform_(1e-9, 1e-7, generator=gen)
        nb = 24
        batches = np.empty((nb, 3, B), np.int32)
        for s_ in range(nb):
            batches[s_, 0] = rng.permutation(np.unique(rng.randint(0, U_n, 2 * B)))[:B]  # distinct users
            batches[s_, 1] = rng.randint(0, I_n, B)
            batches[s_, 2] = rng.randint(0, I_n, B)
        d_b = torch.from_numpy(batches).to(dev)
        losses = torch.zeros((nb, 4), dtype=torch.float32, device=dev)
        tr.run(d_b[:4], losses[:4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.run(d_b[4:], losses[4:])
        e1.record()
        torch.cuda.synchronize()
        t_step = e0.elapsed_time(e1) * 1e-3 / (nb - 4)
        step_bytes = 24.0 * D * (U_n + I_n) + 12.0 * D * B + 12.0 * B + 48.0 * D
        emit(config="MACR-MF %s U=%d I=%d d=64 B=8192 rubibceboth, %.3f GB of var/m/v"
                    % (tname, U_n, I_n, 12.0 * D * (U_n + I_n) / 1e9),
             metric="train_interactions_per_sec", value=B / t_step, unit="interactions/s", ms_per_step=1e3 * t_step,
             roofline={"bound": "hbm", "bytes_per_step": step_bytes, "achieved": step_bytes / t_step / 1e9,
                       "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": step_bytes / t_step / 1e9 / pk["hbm_gbs"],
                       "note": "whole fused step (gather + dots + BxB BCE + row gradients + dense Adam) against "
                               "its algorithmic bytes 24*64*(U+I) + 780*B + 3072"},
             final_loss=float(losses[nb - 1, 0].item()))
        tr.close()
        del tr
        torch.cuda.empty_cache()

    if "ml10m" in what:
        scoring("MACR-MF ml_10m-shape scoring T=13878 I=8790 K=20 c=40", 13878, 8790, 71, 20)
    if "synth" in what:
        scoring("synthetic %d users x %d items d=64 K=20 c=40 (config 5 query slice)" %
                (args.synth_users, args.synth_items), args.synth_users, args.synth_items, 20, 2)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
