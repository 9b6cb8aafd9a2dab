#!/usr/bin/env python
"""Regenerates tests/golden/* by running the REFERENCE's own code (read-only, from
/root/reference) in THIS container.  The reference cannot travel to the GPU box, so the small
outputs are committed as fixtures next to this script.

What is produced (all from reference code, none from this repo's implementation):
  tiny/{train,test}.txt    a synthetic data set in the reference's text format (seeded here)
  mf_sampler.npz           macr_mf/load_data.py Data.sample()      (:543-566) on tiny + addressa digests
  lgcn_sampler.npz         macr_lightgcn/utility/load_data.py Data.sample() (:174-212)
  lgcn_sample_test.npz     Data.sample_test() (:213-257), run with random.sample's pre-3.11 Set handling
  lgcn_adj_tiny.npz        Data.get_adj_mat() `pre` adjacency (:95-124) of tiny, full CSR
  lgcn_adj_variants_tiny.npz  the plain / norm / mean members of the same 4-tuple (:126-164)
  digests.json             sha1 / statistics of the same objects on the real addressa data
  ref_evaluator.npz        the reference's C++ evaluator (oracle/_ref, built from
                           evaluator/cpp/include/*.h) on a seeded score matrix
  parser_defaults.json     defaults of macr_mf/parse.py and macr_lightgcn/utility/parser.py
  lgcn_test.npz            macr_lightgcn/utility/batch_test.py's own test() (:26-162: train items := -inf,
                           the C++ evaluator -- here the reference's, oracle/_ref --, the hit-ratio rewrite
                           of slot 2, the user mean at the cut-offs Ks) on seeded score matrices of `tiny`
  mf_metrics.npz           macr_mf/train.py's own ranklist_by_sorted + get_performance (:32-117) on a
                           seeded score matrix (the functions are compiled from the reference file at
                           generation time -- the module itself imports tensorflow and cannot be loaded)

Usage:  python tests/golden/make_golden.py        (needs /root/reference and oracle/_ref)
"""
import hashlib
import importlib.util
import json
import os
import random
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MACR_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)


def load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def stub_matplotlib():
    """macr_mf/load_data.py:9,14 imports pyplot only to call switch_backend."""
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.switch_backend = lambda *_a, **_k: None
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def write_tiny(path):
    """80 users x 50 items; shuffled user lines; user 7 has no train line; item 49 only in test."""
    rng = np.random.RandomState(2024)
    n_users, n_items = 80, 50
    os.makedirs(path, exist_ok=True)
    order = rng.permutation(n_users)
    train, test = {}, {}
    for u in range(n_users):
        k = int(rng.randint(3, 13))
        items = rng.choice(n_items - 1, size=k + 3, replace=False)
        train[u] = [int(x) for x in items[:k]]
        test[u] = [int(x) for x in items[k:k + int(rng.randint(1, 4))]]
    test[3].append(49)
    with open(os.path.join(path, "train.txt"), "w") as f:
        for u in order:
            if u == 7:
                continue
            f.write(" ".join(str(x) for x in [u] + train[u]) + "\n")
    with open(os.path.join(path, "test.txt"), "w") as f:
        for u in sorted(test):
            if u % 9 == 4:
                continue  # users without test items
            f.write(" ".join(str(x) for x in [u] + test[u]) + "\n")


def sha1_triples(u, p, n):
    return hashlib.sha1(np.array([u, p, n], np.int64).tobytes()).hexdigest()


def mf_args(dataset, batch_size):
    return types.SimpleNamespace(data_path="./data/", dataset=dataset, batch_size=batch_size,
                                 data_type="ori", model="mf", source="normal", valid_set="test")


def run_mf(workdir, digests, out):
    stub_matplotlib()
    mod = load_module("ref_mf_load_data", os.path.join(REF, "macr_mf", "load_data.py"))
    cwd = os.getcwd()
    os.chdir(workdir)  # the MF loader reads ./data/<dataset>/ relative to cwd (load_data.py:27)
    try:
        data = mod.Data(mf_args("tiny", 16))
        random.seed(12345)
        b1 = data.sample()
        b2 = data.sample()
        data.batch_size = 200  # > n_users: the rd.choice branch (load_data.py:546-547)
        b3 = data.sample()
        out["mf_sampler"] = dict(
            n_users=data.n_users, n_items=data.n_items, n_train=data.n_train, n_test=data.n_test,
            b1=np.array(b1, np.int64), b2=np.array(b2, np.int64), b3=np.array(b3, np.int64))
        if os.path.isdir("data/addressa"):
            data = mod.Data(mf_args("addressa", 1024))
            random.seed(12345)
            u, p, n = data.sample()
            digests["mf_addressa"] = dict(
                n_users=data.n_users, n_items=data.n_items, n_train=data.n_train,
                n_test=data.n_test, n_test_users=len(data.test_users), first_batch_sha1=sha1_triples(u, p, n),
                users5=u[:5], pos5=p[:5], neg5=n[:5])
    finally:
        os.chdir(cwd)


def run_lgcn(workdir, digests, out):
    mod = load_module("ref_lgcn_load_data", os.path.join(REF, "macr_lightgcn", "utility", "load_data.py"))
    args = types.SimpleNamespace(valid_set="test")
    data = mod.Data(path=os.path.join(workdir, "data", "tiny"), batch_size=16, args=args)
    random.seed(12345)
    np.random.seed(12345)
    b1 = data.sample()
    b2 = data.sample()
    out["lgcn_sampler"] = dict(n_users=data.n_users, n_items=data.n_items, n_train=data.n_train,
                               n_test=data.n_test, exist_users=np.array(data.exist_users, np.int64),
                               b1=np.array(b1, np.int64), b2=np.array(b2, np.int64))
    # sample_test() (:213-257) calls random.sample on dict keys: a TypeError since Python 3.11.  Up to
    # 3.10 random.sample turned any Set into tuple(population) first; with exactly that conversion
    # restored the reference's own code runs and is pinned here.
    import collections.abc

    orig_sample = random.sample
    random.sample = lambda pop, k: orig_sample(tuple(pop) if isinstance(pop, collections.abc.Set) else pop, k)
    try:
        random.seed(4321)
        np.random.seed(4321)
        t1 = data.sample_test()
        t2 = data.sample_test()
    finally:
        random.sample = orig_sample
    out["lgcn_sample_test"] = dict(t1=np.array(t1, np.int64), t2=np.array(t2, np.int64))
    plain, norm, mean, pre = data.get_adj_mat()
    pre = pre.tocsr()
    pre.sort_indices()
    out["lgcn_adj_tiny"] = dict(indptr=pre.indptr.astype(np.int32), indices=pre.indices.astype(np.int32),
                                data=pre.data.astype(np.float32), shape=np.array(pre.shape))
    # the other members of the 4-tuple (--adj_type plain | norm | gcmc, LightGCN.py:666-681); the
    # reference's `norm` is float64, the session feeds float32 (LightGCN.py:537-540)
    variants = {}
    for name, M in (("plain", plain), ("norm", norm), ("mean", mean)):
        M = M.tocsr()
        M.sort_indices()
        variants.update({name + "_indptr": M.indptr.astype(np.int32), name + "_indices": M.indices.astype(np.int32),
                         name + "_data": M.data.astype(np.float32)})
    out["lgcn_adj_variants_tiny"] = variants
    adir = os.path.join(workdir, "data", "addressa")
    if os.path.isdir(adir):
        data = mod.Data(path=adir, batch_size=1024, args=args)
        random.seed(12345)
        np.random.seed(12345)
        u, p, n = data.sample()
        _, _, _, pre = data.get_adj_mat()
        pre = pre.tocsr()
        pre.sort_indices()
        h = hashlib.sha1()
        for a in (pre.indptr.astype(np.int32), pre.indices.astype(np.int32), pre.data.astype(np.float32)):
            h.update(a.tobytes())
        digests["lgcn_addressa"] = dict(
            n_users=data.n_users, n_items=data.n_items, n_train=data.n_train, n_test=data.n_test,
            first_batch_sha1=sha1_triples(u, [int(x) for x in p], [int(x) for x in n]),
            users5=[int(x) for x in u[:5]], pos5=[int(x) for x in p[:5]], neg5=[int(x) for x in n[:5]],
            adj_shape=list(pre.shape), adj_nnz=int(pre.nnz), adj_dtype=str(pre.dtype),
            adj_data5=[float(x) for x in pre.data[:5]], adj_row0_sum=float(pre[0].sum()),
            adj_sha1=h.hexdigest())


def run_evaluator(out):
    from oracle import ref_eval

    if not ref_eval.available():
        raise SystemExit("oracle/_ref/libmacr_ref_eval.so missing: run `make -C oracle ref`")
    rng = np.random.RandomState(99)
    rows, cols, K = 48, 300, 20
    scores = rng.randn(rows, cols).astype(np.float32)  # tie-free with probability ~1
    truth = [np.sort(rng.choice(cols, size=int(rng.randint(1, 30)), replace=False)).astype(np.int32)
             for _ in range(rows)]
    rk = ref_eval.top_k_array_index(scores, K, thread_num=4)
    res = ref_eval.evaluate_foldout(rk, truth, thread_num=4)
    out["ref_evaluator"] = dict(scores=scores, truth_len=np.array([len(t) for t in truth], np.int32),
                                truth=np.concatenate(truth), rankings=rk, results=res)


def run_lgcn_test(workdir, out):
    """Executes the reference's OWN LightGCN `test()` (utility/batch_test.py:26-162).  The module
    parses flags, loads data and imports the Cython evaluator on import, so `test` is lifted out of
    its syntax tree and compiled as it stands into a namespace holding the reference's own `Data`
    object, the reference's own C++ evaluator (oracle/_ref) and a session stub that returns seeded
    score matrices."""
    import ast
    import heapq

    from oracle import ref_eval

    mod = load_module("ref_lgcn_load_data_t", os.path.join(REF, "macr_lightgcn", "utility", "load_data.py"))
    data = mod.Data(path=os.path.join(workdir, "data", "tiny"), batch_size=16,
                    args=types.SimpleNamespace(valid_set="test"))
    path = os.path.join(REF, "macr_lightgcn", "utility", "batch_test.py")
    tree = ast.parse(open(path).read(), filename=path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "test"]
    assert len(body) == 1
    ns = {"np": np, "heapq": heapq, "data_generator": data, "BATCH_SIZE": 16, "ITEM_NUM": data.n_items,
          "USR_NUM": data.n_users, "args": types.SimpleNamespace(layer_size="[64,64]"),
          "eval_score_matrix_foldout": ref_eval.eval_score_matrix_foldout}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)

    rng = np.random.RandomState(55)
    mats = {"batch_ratings": rng.randn(data.n_users, data.n_items).astype(np.float32),
            "rubi_ratings_both": rng.randn(data.n_users, data.n_items).astype(np.float32)}

    class Model:
        Ks = [20, 5]  # unsorted on purpose: test() sorts them (:30)
        users, pos_items, node_dropout, mess_dropout = "users", "pos_items", "node_dropout", "mess_dropout"
        batch_ratings, rubi_ratings_both = "batch_ratings", "rubi_ratings_both"

    class Sess:
        def run(self, fetch, feed):
            return mats[fetch][np.asarray(feed["users"], dtype=np.int64)].copy()

    users_to_test = list(data.test_set.keys())
    cwd = os.getcwd()
    os.chdir(workdir)  # test() writes Lightgcn_macr.txt into the cwd (:116-117)
    try:
        res = {m: ns["test"](Sess(), Model(), users_to_test, method=m) for m in ("normal", "rubiboth")}
    finally:
        os.chdir(cwd)
    out["lgcn_test"] = dict(users_to_test=np.array(users_to_test, np.int32), Ks=np.array(Model.Ks),
                            batch_ratings=mats["batch_ratings"], rubi_ratings_both=mats["rubi_ratings_both"],
                            **{f"{m}_{k}": np.asarray(v, np.float64) for m, r in res.items() for k, v in r.items()})


def run_mf_metrics(out):
    """Executes the reference's OWN metric functions (macr_mf/train.py:32-117).  train.py imports
    tensorflow at the top, so the function definitions are lifted out of its syntax tree and
    compiled as they stand; `np.asfarray` (removed in numpy 2) is given its numpy-1.x meaning."""
    import ast
    import heapq
    import math

    path = os.path.join(REF, "macr_mf", "train.py")
    tree = ast.parse(open(path).read(), filename=path)
    want = {"precision_at_k", "dcg_at_k", "ndcg_at_k", "recall_at_k", "hit_at_k", "ranklist_by_sorted",
            "get_performance"}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in body} == want
    np_shim = types.ModuleType("numpy_with_asfarray")
    np_shim.__dict__.update(np.__dict__)
    np_shim.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)  # numpy 1.x definition
    ns = {"np": np_shim, "heapq": heapq, "math": math}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)

    rng = np.random.RandomState(77)
    n_users, n_items, Ks = 40, 300, [5, 20]
    rating = rng.randn(n_users, n_items).astype(np.float32)
    rating[3, 10] = rating[3, 200] = 9.0  # one exact tie on top: heapq.nlargest keeps the lower id first
    train = [np.sort(rng.choice(n_items, size=int(rng.randint(0, 40)), replace=False)) for _ in range(n_users)]
    test = []
    for u in range(n_users):
        rest = np.setdiff1d(np.arange(n_items), train[u])
        test.append(np.sort(rng.choice(rest, size=int(rng.randint(1, 12)), replace=False)))
    test[3] = np.array([10])  # the tied pair decides a hit at K = 1..: id 10 must rank before id 200
    result = {k: np.zeros(len(Ks)) for k in ("precision", "recall", "ndcg", "hit_ratio")}
    hits = np.zeros((n_users, max(Ks)), np.int8)
    for u in range(n_users):  # test_one_user (:119-138) + the accumulation of test() (:286-290)
        test_items = list(set(range(n_items)) - set(train[u].tolist()))
        r = ns["ranklist_by_sorted"](test[u].tolist(), test_items, rating[u], Ks)
        hits[u] = r
        re = ns["get_performance"](test[u].tolist(), r, Ks)
        for k in result:
            result[k] += re[k] / n_users
    out["mf_metrics"] = dict(rating=rating, Ks=np.array(Ks), hits=hits,
                             train_len=np.array([len(t) for t in train], np.int32), train=np.concatenate(train).astype(np.int32),
                             test_len=np.array([len(t) for t in test], np.int32), test=np.concatenate(test).astype(np.int32),
                             **{"res_" + k: v for k, v in result.items()})


def run_parsers(digests):
    out = {}
    for key, path in (("mf", "macr_mf/parse.py"), ("lgcn", "macr_lightgcn/utility/parser.py")):
        mod = load_module("ref_parser_" + key, os.path.join(REF, path))
        argv, sys.argv = sys.argv, ["x"]
        try:
            out[key] = vars(mod.parse_args())
        finally:
            sys.argv = argv
    return out


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not mounted: golden fixtures can only be regenerated next to the reference")
    write_tiny(os.path.join(HERE, "tiny"))
    work = tempfile.mkdtemp(prefix="macr_golden_")
    try:
        os.makedirs(os.path.join(work, "data"))
        shutil.copytree(os.path.join(HERE, "tiny"), os.path.join(work, "data", "tiny"))
        src = os.path.join(REF, "data", "addressa")
        if os.path.isdir(src):  # writable copy: get_adj_mat() writes .npz into the data dir
            shutil.copytree(src, os.path.join(work, "data", "addressa"))
        digests, out = {}, {}
        run_mf(work, digests, out)
        run_lgcn(work, digests, out)
        run_evaluator(out)
        run_mf_metrics(out)
        run_lgcn_test(work, out)
        for name, d in out.items():
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        with open(os.path.join(HERE, "digests.json"), "w") as f:
            json.dump(digests, f, indent=1, sort_keys=True)
        with open(os.path.join(HERE, "parser_defaults.json"), "w") as f:
            json.dump(run_parsers(digests), f, indent=1, sort_keys=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
