"""-m gpu: the host-side mirror of the reference interface (session facades, evaluators, CLI
drivers, checkpoints) on the tiny data set of tests/golden, checked against the CPU oracle."""
import os
import random
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def mf_args(**kw):
    from macr_b200.host import flags

    argv = ["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--train", "rubibceboth",
            "--test", "rubi", "--c", "2.0", "--alpha", "1e-2", "--beta", "1e-3", "--lr", "0.01",
            "--save_flag", "0", "--verbose", "0"]
    a = flags.parse_mf_args(argv)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def lgcn_args(**kw):
    from macr_b200.host import flags

    argv = ["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--layer_size", "[64,64]",
            "--loss", "bceboth", "--test", "rubiboth", "--c", "2.0", "--alpha", "1e-2", "--beta", "1e-3",
            "--lr", "0.01", "--Ks", "[5,20]", "--save_flag", "0", "--verbose", "0"]
    a = flags.parse_lgcn_args(argv)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def test_mf_facade_steps_scores_and_eval_match_oracle(oracle):
    from macr_b200.host.data_mf import Data
    from macr_b200.host.evaluate import MFEvaluator
    from macr_b200.host.model_mf import BPRMF, init_weights
    from macr_b200.host.session import Session
    from oracle import mf_metrics

    args = mf_args()
    data = Data(args)
    model = BPRMF(args, {"n_users": data.n_users, "n_items": data.n_items})
    sess = Session()
    U, I, w, wu = init_weights(data.n_users, data.n_items, 64, args.init_seed)
    U, I = U * 8, I * 8  # non-degenerate scores
    model.load_state_dict({**{k: np.zeros_like(v) for k, v in model.state_dict().items() if k[0] in "mv"},
                           "U": U, "I": I, "w": w, "wu": wu, "steps_done": 0})
    st = oracle.MFState(U, I, w, wu)
    hp = oracle.HParams.make(lr=args.lr, alpha=args.alpha, beta=args.beta, decay=args.regs,
                             batch_size=args.batch_size)
    random.seed(12345)
    for _ in range(5):
        users, pos, neg = data.sample()
        got = sess.run([model.opt_two_bce_both, model.loss_two_bce_both, model.mf_loss_two_bce_both,
                        model.reg_loss_two_bce_both],
                       feed_dict={model.users: users, model.pos_items: pos, model.neg_items: neg})
        want = oracle.mf_step(st, users, pos, neg, hp)
        assert got[0] is None
        np.testing.assert_allclose(got[1:], want[:3], rtol=1e-4, atol=1e-4)
    # the literal rubi_ratings_both fetch
    model.update_c(sess, args.c)
    test_users = list(data.test_user_list.keys())
    M = sess.run(model.rubi_ratings_both, {model.users: test_users[:10], model.pos_items: list(range(data.n_items))})
    assert M.shape == (10, data.n_items) and M.dtype == np.float32
    Uq = np.ascontiguousarray(st.U[test_users[:10]])
    want_M = oracle.score_matrix(Uq, st.I, oracle.score_gates(st.I, st.w), oracle.score_gates(Uq, st.wu), args.c)
    np.testing.assert_allclose(M, want_M, rtol=1e-4, atol=1e-5)
    B0 = sess.run(model.batch_ratings, {model.users: test_users[:10], model.pos_items: range(data.n_items)})
    np.testing.assert_allclose(B0, Uq @ st.I.T, rtol=1e-4, atol=1e-5)
    # evaluation: fused == literal matrix path == oracle ranking + reference metric arithmetic
    Ks = [5, 20]
    fused = MFEvaluator(data, Ks, 32, "fused").test(sess, model, test_users, model_type="rubi_both")
    literal = MFEvaluator(data, Ks, 32, "matrix").test(sess, model, test_users, model_type="rubi_both")
    t = model.trainer.tab
    Ud, Id, wd, wud = (x.cpu().numpy() for x in (t.U, t.I, t.w, t.wu))
    Uq = np.ascontiguousarray(Ud[test_users])
    mrp, mcol = data.train_csr(test_users)
    ids, _ = oracle.score_topk(Uq, Id, oracle.score_gates(Id, wd), oracle.score_gates(Uq, wud), args.c,
                               mrp, mcol, 20)
    want = mf_metrics.evaluate(ids, [data.test_user_list[u] for u in test_users], Ks)
    for k in want:
        np.testing.assert_allclose(fused[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)
        np.testing.assert_allclose(literal[k], want[k], rtol=1e-9, atol=1e-12, err_msg=k)
    # the item-gate-only head (`--train rubibce --test rubi` -> model_type rubi_c, train.py:241,551)
    ones = np.ones(len(test_users), np.float32)
    M1 = sess.run(model.rubi_ratings, {model.users: test_users[:10], model.pos_items: list(range(data.n_items))})
    want_M1 = oracle.score_matrix(np.ascontiguousarray(Ud[test_users[:10]]), Id, oracle.score_gates(Id, wd),
                                  ones[:10], args.c)
    np.testing.assert_allclose(M1, want_M1, rtol=1e-4, atol=1e-5)
    ids1, _ = oracle.score_topk(Uq, Id, oracle.score_gates(Id, wd), ones, args.c, mrp, mcol, 20)
    want1 = mf_metrics.evaluate(ids1, [data.test_user_list[u] for u in test_users], Ks)
    fused1 = MFEvaluator(data, Ks, 32, "fused").test(sess, model, test_users, model_type="rubi_c")
    literal1 = MFEvaluator(data, Ks, 32, "matrix").test(sess, model, test_users, model_type="rubi_c")
    for k in want1:
        np.testing.assert_allclose(fused1[k], want1[k], rtol=1e-9, atol=1e-12, err_msg=k)
        np.testing.assert_allclose(literal1[k], want1[k], rtol=1e-9, atol=1e-12, err_msg=k)
    # and its training fetch list (train.py:482-486)
    st2 = oracle.MFState(Ud, Id, wd, wud)
    for k in ("mU", "vU", "mI", "vI", "mw", "vw", "mwu", "vwu"):
        getattr(st2, k)[...] = getattr(t, k).cpu().numpy().reshape(getattr(st2, k).shape)
    st2.t = model.trainer.steps_done
    st2.pw[:] = (hp.beta1 ** (st2.t + 1), hp.beta2 ** (st2.t + 1))
    users, pos, neg = data.sample()
    got = sess.run([model.opt_two_bce, model.loss_two_bce, model.mf_loss_two_bce, model.reg_loss_two_bce],
                   feed_dict={model.users: users, model.pos_items: pos, model.neg_items: neg})
    want = oracle.mf_step_item(st2, users, pos, neg, hp)
    assert got[0] is None
    np.testing.assert_allclose(got[1:], want[:3], rtol=1e-4, atol=1e-4)
    model.close()


def test_mf_checkpoint_resume_is_bit_exact(tmp_path):
    from macr_b200.host import checkpoint
    from macr_b200.host.data_mf import Data
    from macr_b200.host.model_mf import BPRMF

    args = mf_args()
    data = Data(args)
    cfg = {"n_users": data.n_users, "n_items": data.n_items}

    def steps(model, n):
        out = []
        for _ in range(n):
            out.append(model.train_step(*data.sample()))
        return out

    random.seed(7)
    a = BPRMF(args, cfg)
    steps(a, 3)
    path = str(tmp_path / "ck.npz")
    checkpoint.save(path, a)
    tail_a = steps(a, 3)
    b = BPRMF(args, cfg)
    random.seed(123)  # wrong stream on purpose: load() restores it
    checkpoint.load(path, b)
    tail_b = steps(b, 3)
    assert tail_a == tail_b
    np.testing.assert_array_equal(a.trainer.tab.U.cpu().numpy(), b.trainer.tab.U.cpu().numpy())
    np.testing.assert_array_equal(a.trainer.tab.vI.cpu().numpy(), b.trainer.tab.vI.cpu().numpy())
    a.close()
    b.close()


def test_lgcn_facade_and_evaluator_match_oracle(oracle):
    from macr_b200.host.data_lgcn import Data
    from macr_b200.host.evaluate import LGCNEvaluator
    from macr_b200.host.model_lgcn import LightGCN
    from macr_b200.host.session import Session
    from helpers import lists_to_csr

    args = lgcn_args()
    data = Data(GOLD + "/tiny", args.batch_size, args)
    _, _, _, pre = data.get_adj_mat()
    model = LightGCN({"n_users": data.n_users, "n_items": data.n_items, "norm_adj": pre}, None, args=args)
    sess = Session()
    sd = model.state_dict()
    st = oracle.MFState(sd["U"], sd["I"], sd["w"], sd["wu"])
    rowptr, col, val = data.adj_csr("pre")
    hp = oracle.HParams.make(lr=args.lr, alpha=args.alpha, beta=args.beta, decay=1e-5, batch_size=args.batch_size)
    random.seed(12345)
    np.random.seed(12345)
    feed = lambda t: {model.users: t[0], model.pos_items: t[1], model.neg_items: t[2],
                      model.node_dropout: [0.1], model.mess_dropout: [0.1]}
    for _ in range(4):
        tr = data.sample()
        lo_eval = oracle.lgcn_step(st, rowptr, col, val, 2, *tr, hp, train=False)
        got_eval = sess.run([model.loss_two_bce_both, model.mf_loss_two_bce_both, model.emb_loss_two_bce_both], feed(tr))
        np.testing.assert_allclose(got_eval, lo_eval[:3], rtol=1e-4, atol=1e-4)
        lo = oracle.lgcn_step(st, rowptr, col, val, 2, *tr, hp, train=True)
        got = sess.run([model.opt_two_bce_both, model.loss_two_bce_both, model.mf_loss_two_bce_both,
                        model.emb_loss_two_bce_both, model.reg_loss_two_bce_both], feed(tr))
        np.testing.assert_allclose(got[1:4], lo[:3], rtol=1e-4, atol=1e-4)
        assert got[0] is None and np.array_equal(got[4], [0.0])
    model.update_c(sess, args.c)
    users = list(data.test_set.keys())
    fused = LGCNEvaluator(data, 32, "fused").test(sess, model, users, method="rubiboth")
    literal = LGCNEvaluator(data, 32, "matrix").test(sess, model, users, method="rubiboth")
    # oracle route: propagate -> score -> -inf mask -> top-K -> foldout curves -> hr rewrite
    E = oracle.lgcn_propagate(rowptr, col, val, st.U, st.I, 2)
    Ue, Ie = E[:data.n_users], np.ascontiguousarray(E[data.n_users:])
    Uq = np.ascontiguousarray(Ue[users])
    mrp, mcol = data.train_csr(users)
    ids, _ = oracle.score_topk(Uq, Ie, oracle.score_gates(Ie, st.w), oracle.score_gates(Uq, st.wu), args.c,
                               mrp, mcol, 20)
    trp, tcol = lists_to_csr([data.test_set[u] for u in users])
    res = oracle.foldout_metrics(ids, trp, tcol)
    res[:, 40:60] = (res[:, 20:40] != 0)
    final = res.mean(0).reshape(5, 20)[:, np.array([5, 20]) - 1]
    for key, row in (("hr", 2), ("recall", 1), ("ndcg", 3)):
        np.testing.assert_allclose(fused[key], final[row], rtol=1e-5, atol=1e-6, err_msg=key)
        np.testing.assert_allclose(literal[key], final[row], rtol=1e-5, atol=1e-6, err_msg=key)
    model.close()


def test_eval_score_matrix_foldout_drop_in_matches_reference_golden():
    from macr_b200.host.evaluate import eval_score_matrix_foldout

    g = np.load(os.path.join(GOLD, "ref_evaluator.npz"))
    off = np.concatenate([[0], np.cumsum(g["truth_len"])])
    truth = [g["truth"][off[i]:off[i + 1]].tolist() for i in range(len(g["truth_len"]))]
    got = eval_score_matrix_foldout(g["scores"], truth, top_k=20, thread_num=8)
    np.testing.assert_array_equal(got, g["results"])  # the reference's own C++ output, bit for bit
    with pytest.raises(ValueError):
        eval_score_matrix_foldout(g["scores"], truth[:-1])


def test_mf_cli_pipelined_sampler_equals_the_stepwise_loop(tmp_path, monkeypatch, capsys):
    """The MF CLI samples epoch k+1 on a worker thread while epoch k trains: the run must be the
    literal `data.sample()` / `sess.run` loop of train.py:470-499 (MACR_STEPWISE=1), and a
    checkpoint of epoch k must carry the sampler stream as of the end of epoch k's draws."""
    from macr_b200.cli import train_mf
    from macr_b200.host.data_mf import Data

    monkeypatch.chdir(tmp_path)
    argv = ["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "4", "--log_interval", "2",
            "--train", "rubibceboth", "--test", "rubi", "--c", "2", "--lr", "0.01"]
    a = train_mf.main(argv + ["--saveID", "pipe"])
    out_a = [l for l in capsys.readouterr().out.splitlines() if "train==" in l]
    monkeypatch.setenv("MACR_STEPWISE", "1")
    b = train_mf.main(argv + ["--saveID", "step", "--save_flag", "0"])
    out_b = [l for l in capsys.readouterr().out.splitlines() if "train==" in l]
    monkeypatch.delenv("MACR_STEPWISE")
    strip = lambda l: l.split("]: ", 1)[1]  # drop the wall-clock prefix
    assert [strip(l) for l in out_a] == [strip(l) for l in out_b] and len(out_a) >= 2
    assert {k: a[k] for k in ("best_hr", "best_ndcg", "best_recall", "best_epoch")} == \
           {k: b[k] for k in ("best_hr", "best_ndcg", "best_recall", "best_epoch")}
    # the checkpoint written after epoch index 1 holds the stream position after 2 epochs of draws,
    # although the worker had already sampled the third epoch when it was written
    z = np.load("mf_tiny_checkpoint/wd_1e-05_lr_0.01_pipe/1_ckpt.npz")
    args = mf_args()
    data = Data(args)
    random.seed(12345)
    data.sample_epoch(2 * (data.n_train // 32 + 1))
    np.testing.assert_array_equal(z["py_random_key"], np.asarray(random.getstate()[1], np.uint32))


def test_lgcn_cli_epoch_path_equals_the_stepwise_thread_loop(tmp_path, monkeypatch, capsys):
    """The LightGCN CLI draws an epoch's n_batch + 1 train batches (and, at logging epochs, the
    n_batch + 1 test batches) in one native call on a worker thread and runs the epoch / the
    loss-only pass as one call each: it must print what the reference's per-step sampler-thread /
    train-thread loop prints (LightGCN.py:762-819, MACR_STEPWISE=1) and end on the same metrics."""
    from macr_b200.cli import lightgcn

    monkeypatch.chdir(tmp_path)
    argv = ["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "4", "--log_interval", "2",
            "--layer_size", "[64,64]", "--Ks", "[20]", "--loss", "bceboth", "--test", "normal", "--lr", "0.001",
            "--verbose", "1", "--save_flag", "0", "--weights_path", str(tmp_path) + "/"]
    a = lightgcn.main(argv)
    out_a = [l for l in capsys.readouterr().out.splitlines() if "train==" in l or "test==" in l]
    monkeypatch.setenv("MACR_STEPWISE", "1")
    b = lightgcn.main(argv)
    out_b = [l for l in capsys.readouterr().out.splitlines() if "train==" in l or "test==" in l]
    monkeypatch.delenv("MACR_STEPWISE")
    strip = lambda l: l.split("]: ", 1)[1]  # drop the wall-clock prefix
    assert len(out_a) == 4 and [strip(l) for l in out_a] == [strip(l) for l in out_b]
    assert a["best_hr"] == b["best_hr"] and a["best_epoch"] == b["best_epoch"]
    for k in ("recall", "hr", "ndcg"):
        np.testing.assert_array_equal(a["last"][k], b["last"][k])


def test_cli_drivers_run_end_to_end(tmp_path, monkeypatch, capsys):
    from macr_b200.cli import lightgcn, train_mf

    monkeypatch.chdir(tmp_path)
    cfg = train_mf.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "4",
                         "--log_interval", "2", "--train", "rubibceboth", "--test", "rubi", "--c", "2",
                         "--lr", "0.01", "--saveID", "t"])
    out = capsys.readouterr().out
    assert "c:2.00 [" in out and "hit=[" in out and "Epoch 0 [" in out
    assert os.path.exists("mf_tiny_checkpoint/wd_1e-05_lr_0.01_t/3_ckpt.npz")
    assert 0 <= cfg["best_hr"] <= 1
    # the README's baseline command (README.md:30): --train normalbce --test normal
    cfg = train_mf.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "4",
                         "--log_interval", "2", "--train", "normalbce", "--test", "normal", "--lr", "0.01",
                         "--saveID", "n", "--save_flag", "0"])
    out = capsys.readouterr().out
    assert "Epoch 1 [" in out and "hit=[" in out and 0 <= cfg["best_hr"] <= 1
    # tune.py: the c sweep of README.md:101-105 at every evaluation
    cfg = train_mf.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "2",
                         "--log_interval", "2", "--train", "rubibceboth", "--test", "rubi", "--start", "0",
                         "--end", "4", "--step", "3", "--lr", "0.01", "--saveID", "u", "--save_flag", "0"],
                        tune=True)
    out = capsys.readouterr().out
    assert "c:0.00 [" in out and "c:2.00 [" in out and "c:4.00 [" in out and "best c:" in out
    assert cfg["best_c"] in (0.0, 2.0, 4.0)
    res = lightgcn.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "2",
                         "--log_interval", "1", "--layer_size", "[64,64]", "--Ks", "[20]", "--loss", "bceboth",
                         "--test", "rubiboth", "--c", "2", "--lr", "0.001", "--weights_path", str(tmp_path) + "/"])
    out = capsys.readouterr().out
    assert "c:2.00 recall=[" in out and "use the pre adjcency matrix" in out
    assert res["last"] is not None and 0 <= res["best_hr"] <= 1
    # LightGCN_tune.py: c sweep at every evaluation
    lightgcn.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "1",
                   "--log_interval", "1", "--layer_size", "[64,64]", "--Ks", "[20]", "--loss", "bceboth",
                   "--test", "rubiboth", "--start", "0", "--end", "4", "--step", "3", "--lr", "0.001",
                   "--weights_path", str(tmp_path) + "/"], tune=True)
    out = capsys.readouterr().out
    assert "c:0.00 recall=[" in out and "c:2.00 recall=[" in out and "c:4.00 recall=[" in out
    # the README's LightGCN baseline (README.md:59): --loss bce --test normal
    res = lightgcn.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "2",
                         "--log_interval", "1", "--layer_size", "[64,64]", "--Ks", "[20]", "--loss", "bce",
                         "--test", "normal", "--lr", "0.001", "--weights_path", str(tmp_path) + "/"])
    out = capsys.readouterr().out
    assert "recall=[" in out and res["last"] is not None
    # the item-gate-only neighbours: MF `--train rubibce --test rubi` (head rubi_c), LightGCN `--loss bce1 --test rubi1`
    cfg = train_mf.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "2",
                         "--log_interval", "2", "--train", "rubibce", "--test", "rubi", "--c", "2", "--lr", "0.01",
                         "--saveID", "i", "--save_flag", "0"])
    out = capsys.readouterr().out
    assert "c:2.00 [" in out and 0 <= cfg["best_hr"] <= 1
    res = lightgcn.main(["--data_path", GOLD + "/", "--dataset", "tiny", "--batch_size", "32", "--epoch", "1",
                         "--log_interval", "1", "--layer_size", "[64,64]", "--Ks", "[20]", "--loss", "bce1",
                         "--test", "rubi1", "--c", "2", "--lr", "0.001", "--weights_path", str(tmp_path) + "/"])
    out = capsys.readouterr().out
    assert "c:2.00 recall=[" in out and res["last"] is not None


def test_epoch_run_host_equals_device_run():
    """`macr_mf_trainer_run_host` (host batches, one H2D / D2H) == `macr_mf_trainer_run`."""
    import torch

    from helpers import make_batch, make_model
    from macr_b200 import ops

    n_users, n_items, B, n = 900, 500, 256, 5
    U, I, w, wu = make_model(31, n_users, n_items, scale=4.0)
    rng = np.random.RandomState(32)
    batches = np.stack([np.stack(make_batch(rng, n_users, n_items, B)) for _ in range(n)]).astype(np.int32)
    hp = ops.HParams.make(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-5, batch_size=B)
    a = ops.MFTrainer(U, I, w, wu, hp, max_batch=B)
    b = ops.MFTrainer(U, I, w, wu, hp, max_batch=B)
    la = a.run(torch.from_numpy(batches).cuda()).cpu().numpy()
    lb = b.run_host(torch.from_numpy(batches).pin_memory()).numpy()
    np.testing.assert_array_equal(la, lb)
    for x, y in zip(a.tab.all(), b.tab.all()):
        assert torch.equal(x, y)
    assert b.steps_done == n
    a.close()
    b.close()


def test_epoch_path_equals_stepwise_session_loop():
    """`Data.sample_epoch` + `BPRMF.train_epoch` (what the CLI runs) == the literal loop of
    train.py:470-499 (`data.sample()` + `sess.run([...opt...])` per step): losses and tables."""
    import torch

    from macr_b200.host.data_mf import Data
    from macr_b200.host.model_mf import BPRMF
    from macr_b200.host.session import Session

    args = mf_args()
    data = Data(args)
    cfg = {"n_users": data.n_users, "n_items": data.n_items}
    n = 6
    random.seed(21)
    a = BPRMF(args, cfg)
    sess = Session()
    want = []
    for _ in range(n):
        u, p, ng = data.sample()
        _, l, m, r = sess.run([a.opt_two_bce_both, a.loss_two_bce_both, a.mf_loss_two_bce_both,
                               a.reg_loss_two_bce_both],
                              feed_dict={a.users: u, a.pos_items: p, a.neg_items: ng})
        want.append((l, m, r))
    random.seed(21)
    b = BPRMF(args, cfg)
    got = b.train_epoch(data.sample_epoch(n))
    assert [tuple(float(x) for x in row[:3]) for row in got] == want
    for x, y in zip(a.trainer.tab.all(), b.trainer.tab.all()):
        assert torch.equal(x, y)
    a.close()
    b.close()


def test_facade_topk_with_an_evaluation_cache_equals_the_plain_call():
    """`model.topk(..., prep=cache)` (what both evaluators pass: item gates and tensor-core item
    operands prepared by the first user batch of an evaluation, reused by the others) returns the
    ids and scores of the plain call bit for bit -- tensor-core catalogue (5000 items), all three
    score heads; a parameter change between evaluations is picked up by a fresh cache."""
    import torch

    from macr_b200.host.model_mf import BPRMF

    n_users, n_items, K = 300, 5000, 20
    model = BPRMF(mf_args(), {"n_users": n_users, "n_items": n_items})
    with torch.no_grad():
        model.trainer.tab.U.mul_(8.0), model.trainer.tab.I.mul_(8.0)
    model.update_c(None, 30.0)
    rng = np.random.RandomState(2)
    batches = [list(range(0, 130)), list(range(130, 131)), list(range(131, 300))]
    masks = []
    for b in batches:
        rows = [np.sort(rng.choice(n_items, size=rng.randint(0, 40), replace=False)) for _ in b]
        rp = np.zeros(len(b) + 1, np.int32)
        rp[1:] = np.cumsum([len(r) for r in rows])
        masks.append((rp, np.concatenate(rows).astype(np.int32) if rp[-1] else np.zeros(0, np.int32)))
    for head in ("both", "item", "plain"):
        cache = {}
        for b, (rp, col) in zip(batches, masks):
            want_i, want_s = model.topk(b, K, rp, col, head=head)
            got_i, got_s = model.topk(b, K, rp, col, head=head, prep=cache)
            assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
        assert len(cache) == 1
    with torch.no_grad():
        model.trainer.tab.I.mul_(0.5)
    want_i, _ = model.topk(batches[0], K, *masks[0])
    got_i, _ = model.topk(batches[0], K, *masks[0], prep={})
    assert torch.equal(got_i, want_i)
    model.close()
