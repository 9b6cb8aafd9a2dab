"""`python macr_lightgcn/LightGCN.py ...` -- the LightGCN driver of the reference
(macr_lightgcn/LightGCN.py:649-905) on the B200 path: same flags, the same double-buffered
sampler / train threads, the same test-loss pass and early stopping, same log lines.
Only `--alg_type lightgcn --loss bceboth` (with `--test rubiboth` or `--test normal`) is
implemented."""
import logging
from concurrent.futures import ThreadPoolExecutor
import os
import random
import sys
import threading
from time import time

import numpy as np
import scipy.sparse as sp

from ..host import checkpoint, flags
from ..host.data_lgcn import Data
from ..host.evaluate import LGCNEvaluator
from ..host.model_lgcn import LightGCN
from ..host.session import Session


def early_stopping(log_value, best_value, stopping_step, expected_order="acc", flag_step=100):
    """utility/helper.py:35-50."""
    assert expected_order in ("acc", "dec")
    better = log_value >= best_value if expected_order == "acc" else log_value <= best_value
    if better:
        stopping_step, best_value = 0, log_value
    else:
        stopping_step += 1
    should_stop = stopping_step >= flag_step
    if should_stop:
        print("Early stopping is trigger at step: {} log:{}".format(flag_step, log_value))
    return best_value, stopping_step, should_stop


class _Worker(threading.Thread):
    """sample_thread / train_thread of LightGCN.py:566-647: run `fn`, keep its result in .data."""

    def __init__(self, fn):
        super().__init__()
        self.fn, self.data, self.error = fn, None, None

    def run(self):
        try:
            self.data = self.fn()
        except BaseException as e:  # surfaced by the joining thread
            self.error = e

    def result(self):
        self.join()
        if self.error is not None:
            raise self.error
        return self.data


def main(argv=None, tune=False):
    """tune=True is `python macr_lightgcn/LightGCN_tune.py ...` (LightGCN_tune.py:848-871): every
    evaluation sweeps c over np.linspace(--start, --end, --step); the propagated embeddings are
    computed once per parameter version, so a sweep costs one scoring pass per c."""
    args = flags.parse_lgcn_args(argv)
    logging.getLogger().setLevel(logging.INFO)
    if args.alg_type != "lightgcn" or args.loss not in ("bceboth", "bce", "bce1"):
        raise SystemExit(f"--alg_type {args.alg_type} --loss {args.loss}: only lightgcn with bceboth (MACR), bce1 "
                         "(item gate only) or bce (the README's baseline) is implemented on the B200 path "
                         "(DESIGN.md section 8)")
    flags.check_device_limits(args.batch_size, flags.as_list(args.Ks))
    data_generator = Data(path=args.data_path + args.dataset, batch_size=args.batch_size, args=args)
    seed = 12345  # LightGCN.py:651-655
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    logging.basicConfig(filename="LightGCN_{}_{}_{}_{}".format(args.dataset, args.loss, args.test, args.alpha))
    config = {"n_users": data_generator.n_users, "n_items": data_generator.n_items}
    plain_adj, norm_adj, mean_adj, pre_adj = data_generator.get_adj_mat()
    if args.adj_type == "plain":
        config["norm_adj"] = plain_adj
        print("use the plain adjacency matrix")
    elif args.adj_type == "norm":
        config["norm_adj"] = norm_adj
        print("use the normalized adjacency matrix")
    elif args.adj_type == "gcmc":
        config["norm_adj"] = mean_adj
        print("use the gcmc adjacency matrix")
    elif args.adj_type == "pre":
        config["norm_adj"] = pre_adj
        print("use the pre adjcency matrix")
    else:
        config["norm_adj"] = mean_adj + sp.eye(mean_adj.shape[0])
        print("use the mean adjacency matrix")
    pretrain_data = None
    if args.pretrain == -1:  # load_pretrained_data, LightGCN.py:557-564
        path = "%spretrain/%s/%s.npz" % (args.proj_path, args.dataset, "embedding")
        pretrain_data = np.load(path)
        print("load the pretrained embeddings.")
    model = LightGCN(data_config=config, pretrain_data=pretrain_data, args=args)
    sess = Session()
    evaluator = LGCNEvaluator(data_generator, args.batch_size, eval_mode=args.eval_mode)
    layer = "-".join(str(l) for l in flags.as_list(args.layer_size))
    weights_save_path = "%sweights/%s/%s/%s/l%s_r%s" % (
        args.weights_path, args.dataset, model.model_type, layer, str(args.lr),
        "-".join(str(r) for r in flags.as_list(args.regs)))
    if args.save_flag == 1:
        os.makedirs(weights_save_path, exist_ok=True)

    if args.pretrain not in (0, -1, 1):
        raise SystemExit(f"--pretrain {args.pretrain}: only 0, -1 (pretrained embeddings) and 1 (restore + test) exist")
    if args.pretrain == 1:
        # LightGCN.py:707-719 restores a hard-coded "Your Path" and tests c in [0, best_c]; here the
        # file is $MACR_MODEL_FILE or the newest weights_<saveID>-<epoch>.npz under the save path
        import glob

        model_file = os.environ.get("MACR_MODEL_FILE") or next(iter(sorted(
            glob.glob(weights_save_path + "/weights_{}-*.npz".format(args.saveID)),
            key=lambda f: int(f.rsplit("-", 1)[1][:-4]), reverse=True)), None)
        if not model_file:
            raise SystemExit("--pretrain 1: no checkpoint (set MACR_MODEL_FILE or train with --save_flag 1)")
        checkpoint.load(model_file, model, restore_rng=False)
        users_to_test = list(data_generator.test_set.keys())
        ret = None
        for c in [0, args.c]:
            model.update_c(sess, c)
            ret = evaluator.test(sess, model, users_to_test, method="rubiboth")
            print("c:{}: recall={}, hit={}, ndcg={}".format(c, str(ret["recall"]), str(ret["hr"]), str(ret["ndcg"])))
        model.close()
        return {"last": ret, "model_file": model_file}
    if args.only_test != 0:  # LightGCN.py:752 gates the whole training loop on only_test == 0
        print("--only_test %d: training loop skipped (LightGCN.py:752)" % args.only_test)
        model.close()
        return {"best_epoch": 0, "best_hr": 0.0, "last": None}

    if args.loss == "bce":  # LightGCN.py:592-593,625-626
        train_fetch = [model.opt_bce, model.loss_bce, model.mf_loss_bce, model.emb_loss_bce, model.reg_loss_bce]
        test_fetch = [model.loss_bce, model.mf_loss_bce, model.emb_loss_bce]
    elif args.loss == "bce1":  # LightGCN.py:594-595,627-628
        train_fetch = [model.opt_two_bce1, model.loss_two_bce1, model.mf_loss_two_bce1, model.emb_loss_two_bce1,
                       model.reg_loss_two_bce1]
        test_fetch = [model.loss_two_bce1, model.mf_loss_two_bce1, model.emb_loss_two_bce1]
    else:
        train_fetch = [model.opt_two_bce_both, model.loss_two_bce_both, model.mf_loss_two_bce_both,
                       model.emb_loss_two_bce_both, model.reg_loss_two_bce_both]
        test_fetch = [model.loss_two_bce_both, model.mf_loss_two_bce_both, model.emb_loss_two_bce_both]
    drops = {model.node_dropout: flags.as_list(args.node_dropout),
             model.mess_dropout: flags.as_list(args.mess_dropout)}

    def run_on(fetch, triple):
        users, pos_items, neg_items = triple
        feed = {model.users: users, model.pos_items: pos_items, model.neg_items: neg_items}
        feed.update(drops)
        return sess.run(fetch, feed_dict=feed)

    cur_best_pre_0, stopping_step, should_stop = 0.0, 0, False
    best_epoch, ret = 0, None
    config["best_c_hr"], config["best_c_epoch"] = 0, 0
    n_batch = data_generator.n_train // args.batch_size + 1
    stepwise = os.environ.get("MACR_STEPWISE") == "1"
    # Epoch path (default): the reference draws n_batch + 1 train batches per epoch (the last one is
    # sampled by the look-ahead thread and never used, LightGCN.py:762-777) and, at logging epochs,
    # n_batch + 1 test batches -- in that order, from streams nothing else touches.  One native call
    # draws them all (worker thread, one epoch ahead of the device), one call trains the epoch, one
    # call runs the loss-only pass: same triples, same steps, same loss sums as the per-step loop
    # (MACR_STEPWISE=1 keeps that loop: one sampler thread + one train thread per step).
    sampler_pool = None if stepwise else ThreadPoolExecutor(max_workers=1)

    def sample_job(ep):
        tr = data_generator.sample_epoch(n_batch + 1)
        te = data_generator.sample_test_epoch(n_batch + 1) if ep % args.log_interval == 0 else None
        return tr, te, (random.getstate(), np.random.get_state())

    pending = None if stepwise else sampler_pool.submit(sample_job, 1)
    rng_after, test_batches = None, None
    for epoch in range(1, args.epoch + 1):
        t1 = time()
        loss, mf_loss, emb_loss, reg_loss = 0.0, 0.0, 0.0, 0.0
        if not stepwise:
            train_batches, test_batches, rng_after = pending.result()
            pending = sampler_pool.submit(sample_job, epoch + 1) if epoch < args.epoch else None
            for bl in model.run_epoch(train_batches[:n_batch], args.loss, train=True):
                loss += float(bl[0]) / n_batch
                mf_loss += float(bl[1]) / n_batch
                emb_loss += float(bl[2]) / n_batch
        else:
            sample_last = _Worker(data_generator.sample)
            sample_last.start()
            sample_last.join()
            for _ in range(n_batch):  # sampler for step t+1 overlaps the device step t (:762-777)
                triple = sample_last.result()
                train_cur = _Worker(lambda tr=triple: run_on(train_fetch, tr))
                sample_next = _Worker(data_generator.sample)
                train_cur.start()
                sample_next.start()
                sample_next.join()
                _, batch_loss, batch_mf_loss, batch_emb_loss, _ = train_cur.result()
                sample_last = sample_next
                loss += batch_loss / n_batch
                mf_loss += batch_mf_loss / n_batch
                emb_loss += batch_emb_loss / n_batch
        if np.isnan(loss):
            print("ERROR: loss is nan.")
            sys.exit()
        if (epoch % args.log_interval) != 0:
            if args.verbose > 0 and epoch % args.verbose == 0:
                perf_str = "Epoch %d [%.1fs]: train==[%.5f=%.5f + %.5f]" % (epoch, time() - t1, loss, mf_loss, emb_loss)
                print(perf_str)
                logging.info(perf_str)
            continue

        # test loss: n_batch loss-only steps on sample_test() triples (:799-819; consumes RNG)
        loss_test, mf_loss_test, emb_loss_test, reg_loss_test = 0.0, 0.0, 0.0, 0.0
        if not stepwise:
            for bl in model.run_epoch(test_batches[:n_batch], args.loss, train=False):
                loss_test += float(bl[0]) / n_batch
                mf_loss_test += float(bl[1]) / n_batch
                emb_loss_test += float(bl[2]) / n_batch
        else:
            sample_last = _Worker(data_generator.sample_test)
            sample_last.start()
            sample_last.join()
            for _ in range(n_batch):
                triple = sample_last.result()
                train_cur = _Worker(lambda tr=triple: run_on(test_fetch, tr))
                sample_next = _Worker(data_generator.sample_test)
                train_cur.start()
                sample_next.start()
                sample_next.join()
                bl, bm, be = train_cur.result()
                sample_last = sample_next
                loss_test += bl / n_batch
                mf_loss_test += bm / n_batch
                emb_loss_test += be / n_batch

        t2 = time()
        users_to_test = list(data_generator.test_set.keys())
        perf_str = ""
        if args.test == "normal":
            ret = evaluator.test(sess, model, users_to_test, drop_flag=True)
            t3 = time()
            if args.verbose > 0:
                perf_str = ("Epoch %d [%.1fs + %.1fs]: test==[%.5f=%.5f + %.5f + %.5f], recall=[%s], "
                            "hr=[%s], ndcg=[%s]\n") % (
                    epoch, t2 - t1, t3 - t2, loss_test, mf_loss_test, emb_loss_test, reg_loss_test,
                    ", ".join("%.5f" % r for r in ret["recall"]),
                    ", ".join("%.5f" % r for r in ret["hr"]),
                    ", ".join("%.5f" % r for r in ret["ndcg"]))
                print(perf_str, end="")
                logging.info(perf_str)
        elif args.test in ("rubiboth", "rubi1") and tune:
            print("Epoch %d" % epoch)
            best = None
            for c in np.linspace(args.start, args.end, args.step):
                model.update_c(sess, c)
                r = evaluator.test(sess, model, users_to_test, method=args.test)
                if best is None or r["hr"][0] > best[1]["hr"][0]:
                    best = (float(c), r)
                if args.verbose > 0:
                    perf_str += "c:%.2f recall=[%.5f, %.5f], hit=[%.5f, %.5f], ndcg=[%.5f, %.5f]\n" % (
                        c, r["recall"][0], r["recall"][-1], r["hr"][0], r["hr"][-1],
                        r["ndcg"][0], r["ndcg"][-1])
            ret = best[1]
            if ret["hr"][0] > config["best_c_hr"]:
                config.update(best_c_hr=ret["hr"][0], best_c_epoch=epoch, best_c=best[0])
            print(perf_str, end="")
            logging.info(perf_str)
        elif args.test in ("rubiboth", "rubi1"):  # LightGCN.py:848
            print("Epoch %d" % epoch)
            c = args.c
            model.update_c(sess, c)
            ret = evaluator.test(sess, model, users_to_test, method=args.test)
            if args.verbose > 0:
                perf_str += "c:%.2f recall=[%.5f, %.5f], hit=[%.5f, %.5f], ndcg=[%.5f, %.5f]\n" % (
                    c, ret["recall"][0], ret["recall"][-1], ret["hr"][0], ret["hr"][-1],
                    ret["ndcg"][0], ret["ndcg"][-1])
            print(perf_str, end="")
            logging.info(perf_str)
        else:
            raise SystemExit(f"--test {args.test}: only rubiboth / rubi1 / normal are implemented")

        cur_best_pre_0, stopping_step, should_stop = early_stopping(
            ret["hr"][0], cur_best_pre_0, stopping_step, expected_order="acc", flag_step=10)
        if ret["hr"][0] == cur_best_pre_0:
            best_epoch = epoch
        if args.save_flag == 1:
            checkpoint.save(weights_save_path + "/weights_{}-{}.npz".format(args.saveID, epoch), model,
                            {"epoch": epoch}, rng_state=rng_after)
            print("save the weights in path: ", weights_save_path)
        if should_stop and args.early_stop == 1:
            if args.save_flag == 1:
                with open(weights_save_path + "/best_epoch_{}.txt".format(args.saveID), "w") as f:
                    f.write(str(config["best_c_epoch"] if args.test != "normal" else best_epoch))
            break
    if pending is not None:  # early stop: let the speculative draw of the next epoch finish
        pending.result()
    if sampler_pool is not None:
        sampler_pool.shutdown()
    model.close()
    return {"best_epoch": best_epoch, "best_hr": cur_best_pre_0, "last": ret}


if __name__ == "__main__":
    main()
