"""Developer check of the tcgen05 scoring path against the exact fp32 kernel (GPU box only)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from macr_b200 import ops
from macr_b200._lib import lib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_model, make_interactions, lists_to_csr


def run(T_users, n_items, K, c, mask_deg, splits, scale=10.0, seed=0, time_it=False):
    U, I, w, wu = make_model(seed, T_users, n_items, scale=scale)
    lists = make_interactions(seed + 1, T_users, n_items, mask_deg) if mask_deg else None
    dev = torch.device("cuda")
    dU, dI = torch.from_numpy(U).to(dev), torch.from_numpy(I).to(dev)
    si, su = ops.score_gates(dI, torch.from_numpy(w).to(dev)), ops.score_gates(dU, torch.from_numpy(wu).to(dev))
    mrp = mcol = None
    if lists is not None:
        a, b = lists_to_csr(lists)
        mrp, mcol = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    ei, es = ops.score_topk_exact(dU, dI, si, su, c, mrp, mcol, K)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    ti, ts = ops.score_topk_tc(dU, dI, si, su, c, mrp, mcol, K, stats=stats)
    torch.cuda.synchronize()
    bad_rows = int((ti != ei).any(dim=1).sum().item())
    bad_sc = int((ts != es).any(dim=1).sum().item())
    st = stats.cpu().tolist()
    print(f"T={T_users} I={n_items} K={K} c={c} mask={mask_deg} splits={splits}: id-mismatch rows {bad_rows}, "
          f"score-mismatch rows {bad_sc}, fallback rows {st[0]}, candidates/row {st[1] / max(1, T_users - st[0]):.1f}",
          flush=True)
    if bad_rows:
        r = int((ti != ei).any(dim=1).nonzero()[0].item())
        print(" first bad row", r, "\n  tc   ", ti[r].tolist(), "\n  exact", ei[r].tolist())
        print("  tc sc", ts[r].tolist()[:6], "\n  ex sc", es[r].tolist()[:6])
    if time_it:
        for fn, name in ((lambda: ops.score_topk_tc(dU, dI, si, su, c, mrp, mcol, K), "tc"),
                         (lambda: ops.score_topk_exact(dU, dI, si, su, c, mrp, mcol, K), "exact")):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"   {name}: {ms:.3f} ms  {T_users * n_items / ms / 1e6:.1f} G scores/s", flush=True)
    return bad_rows == 0 and bad_sc == 0


def run_trained_like():
    """large norms, low-rank + popularity structure (what trained tables look like), c sweep"""
    dev = torch.device("cuda")
    rng = np.random.RandomState(3)
    T_users, n_items, K = 4096, 40981, 20
    ok = True
    for scale, rank in ((1.0, 8), (3.0, 16), (0.3, 64)):
        Z = rng.randn(T_users, rank).astype(np.float32)
        Wm = rng.randn(rank, 64).astype(np.float32)
        U = (Z @ Wm * scale / np.sqrt(rank)).astype(np.float32)
        pop = rng.pareto(1.2, n_items).astype(np.float32)
        I = ((rng.randn(n_items, rank).astype(np.float32) @ Wm) * scale / np.sqrt(rank) *
             (1.0 + 0.2 * np.log1p(pop))[:, None]).astype(np.float32)
        w = (rng.randn(64) * 0.1).astype(np.float32)
        wu = (rng.randn(64) * 0.1).astype(np.float32)
        lists = make_interactions(11, T_users, n_items, 40)
        a, b = lists_to_csr(lists)
        dU, dI = torch.from_numpy(U).to(dev), torch.from_numpy(I).to(dev)
        si, su = ops.score_gates(dI, torch.from_numpy(w).to(dev)), ops.score_gates(dU, torch.from_numpy(wu).to(dev))
        mrp, mcol = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        for c in (40.0, 0.0, 10.0):
            ei, es = ops.score_topk_exact(dU, dI, si, su, c, mrp, mcol, K)
            st = torch.zeros(2, dtype=torch.int64, device=dev)
            ti, ts = ops.score_topk_tc(dU, dI, si, su, c, mrp, mcol, K, stats=st)
            good = bool((ti == ei).all().item() and (ts == es).all().item())
            print(f"trained-like scale={scale} rank={rank} c={c}: {'ok' if good else 'MISMATCH'} "
                  f"|u|~{np.linalg.norm(U, axis=1).mean():.2f} |i|~{np.linalg.norm(I, axis=1).mean():.2f} "
                  f"fallback {st[0].item()} cand/row {st[1].item() / max(1, T_users - st[0].item()):.1f}", flush=True)
            ok &= good
    return ok


def run_special():
    """heavy train lists, all-equal scores (overflow -> exact fallback), sharded ids, negative c"""
    dev = torch.device("cuda")
    ok = True
    # heavy users: 3000 of 9000 items masked for some rows
    T_users, n_items, K = 200, 9000, 20
    U, I, w, wu = make_model(5, T_users, n_items, scale=10.0)
    rng = np.random.RandomState(5)
    lists = [np.sort(rng.choice(n_items, size=(3000 if u % 7 == 0 else 40), replace=False)).astype(np.int32)
             for u in range(T_users)]
    a, b = lists_to_csr(lists)
    dU, dI = torch.from_numpy(U).to(dev), torch.from_numpy(I).to(dev)
    si, su = ops.score_gates(dI, torch.from_numpy(w).to(dev)), ops.score_gates(dU, torch.from_numpy(wu).to(dev))
    mrp, mcol = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    for c in (40.0, -3.0, 0.0):
        ei, es = ops.score_topk_exact(dU, dI, si, su, c, mrp, mcol, K)
        st = torch.zeros(2, dtype=torch.int64, device=dev)
        ti, ts = ops.score_topk_tc(dU, dI, si, su, c, mrp, mcol, K, stats=st)
        good = bool((ti == ei).all().item() and (ts == es).all().item())
        print(f"heavy masks c={c}: {'ok' if good else 'MISMATCH'} fallback {st[0].item()} cand/row {st[1].item() / T_users:.1f}", flush=True)
        ok &= good
    # sharded: item_id_offset
    lo = 4000
    ei, es = ops.score_topk_exact(dU, dI[lo:].contiguous(), si[lo:].contiguous(), su, 40.0, mrp, mcol, K, item_id_offset=lo)
    ti, ts = ops.score_topk_tc(dU, dI[lo:].contiguous(), si[lo:].contiguous(), su, 40.0, mrp, mcol, K, item_id_offset=lo)
    good = bool((ti == ei).all().item() and (ts == es).all().item())
    print("shard offset:", "ok" if good else "MISMATCH", flush=True)
    ok &= good
    # degenerate: all scores equal -> every item is a candidate -> overflow -> exact fallback
    Z = torch.zeros((150, 64), device=dev)
    ZI = torch.zeros((5000, 64), device=dev)
    h = torch.full((5000,), 0.5, device=dev)
    hu = torch.full((150,), 0.5, device=dev)
    ei, es = ops.score_topk_exact(Z, ZI, h, hu, 40.0, None, None, K)
    st = torch.zeros(2, dtype=torch.int64, device=dev)
    ti, ts = ops.score_topk_tc(Z, ZI, h, hu, 40.0, None, None, K, stats=st)
    good = bool((ti == ei).all().item() and (ts == es).all().item())
    print("all-equal scores:", "ok" if good else "MISMATCH", "fallback rows", st[0].item(), ti[0, :5].tolist(), flush=True)
    ok &= good
    # tiny init (epoch-0 evaluation): scores differ in the last bits
    U, I, w, wu = make_model(9, 300, 6000, scale=1.0)
    dU, dI = torch.from_numpy(U).to(dev), torch.from_numpy(I).to(dev)
    si, su = ops.score_gates(dI, torch.from_numpy(w).to(dev)), ops.score_gates(dU, torch.from_numpy(wu).to(dev))
    for sp in ((1, 1), (1, 3), (3, 3)):
        ei, es = ops.score_topk_exact(dU, dI, si, su, 40.0, None, None, K)
        st = torch.zeros(2, dtype=torch.int64, device=dev)
        ti, ts = ops.score_topk_tc(dU, dI, si, su, 40.0, None, None, K, stats=st)
        good = bool((ti == ei).all().item() and (ts == es).all().item())
        print(f"xavier-init scale=1 splits={sp}:", "ok" if good else "MISMATCH", "fallback", st[0].item(),
              f"cand/row {st[1].item() / max(1, 300 - st[0].item()):.1f}", flush=True)
        ok &= good
    return ok


if __name__ == "__main__":
    ok = True
    ok &= run(128, 5120, 20, 40.0, 0, (1, 1))
    ok &= run(128, 5120, 20, 40.0, 0, (3, 3))
    ok &= run(300, 6000, 20, 40.0, 30, (1, 3))
    ok &= run(300, 6000, 20, 0.0, 30, (3, 3))
    ok &= run(1000, 40981, 20, 40.0, 27, (1, 3))
    ok &= run(15424, 40981, 20, 40.0, 27, (1, 3), time_it=True)
    ok &= run(15424, 40981, 20, 40.0, 27, (1, 1), time_it=True)
    ok &= run(15424, 40981, 20, 40.0, 27, (3, 3), time_it=True)
    ok &= run_special()
    ok &= run_trained_like()
    print("ALL OK" if ok else "FAILURES")
