"""`python ./macr_mf/train.py ...` -- the MF driver of the reference (macr_mf/train.py:332-611)
on the B200 path: same flags, same epoch / evaluation / early-stopping structure, same log
lines.  `--train rubibceboth` (MACR, with `--test rubi` or `--test normal`) and `--train normalbce`
(the README's baseline command, `--test normal`) are implemented; the other --train modes of the
reference are outside the MACR hot path and fail loudly."""
import logging
from concurrent.futures import ThreadPoolExecutor
import os
import random
import sys
from time import time

import numpy as np

from ..host import checkpoint, flags
from ..host.data_mf import Data
from ..host.evaluate import MFEvaluator
from ..host.model_mf import BPRMF
from ..host.session import Session, global_variables_initializer


def early_stop(hr, ndcg, recall, precision, cur_epoch, config, stopping_step, flag_step=10):
    """train.py:313-330: patience `flag_step` evaluations on HR@Ks[0]."""
    if hr >= config["best_hr"]:
        stopping_step = 0
        config.update(best_hr=hr, best_ndcg=ndcg, best_recall=recall, best_pre=precision,
                      best_epoch=cur_epoch)
    else:
        stopping_step += 1
    should_stop = stopping_step >= flag_step
    if should_stop:
        print("Early stopping is trigger")
    return config, stopping_step, should_stop


def ckpt_dir(args):
    return "{}_{}_checkpoint/wd_{}_lr_{}_{}/".format(args.model, args.dataset, args.wd, args.lr, args.saveID)


def main(argv=None, tune=False):
    """tune=True is `python ./macr_mf/tune.py ...` (macr_mf/tune.py:536-575, README.md:101-105):
    the same run, but every evaluation sweeps c over np.linspace(--start, --end, --step) and
    keeps the c with the best HR@Ks[0]."""
    args = flags.parse_mf_args(argv)
    logging.getLogger().setLevel(logging.INFO)
    if args.train not in ("rubibceboth", "normalbce", "rubibce"):
        raise SystemExit(f"--train {args.train}: rubibceboth (MACR), rubibce (item gate only) and normalbce "
                         "(the README's baseline) are implemented on the B200 path (DESIGN.md section 8)")
    if args.model != "mf":
        raise SystemExit(f"--model {args.model}: only mf is on the MACR hot path")
    data = Data(args)
    Ks = flags.as_list(args.Ks)
    flags.check_device_limits(args.batch_size, Ks)
    seed = 12345  # train.py:333-337 (after Data() is built; Data.__init__ draws nothing)
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    logging.basicConfig(filename="{}_{}_{}_{}".format(args.model, args.dataset, args.train, args.wd))
    config = {"n_users": data.n_users, "n_items": data.n_items}
    model = BPRMF(args, config)
    print("MF model.")
    model.set_train_mode(args.train)
    opt_fetches = {  # train.py:482-496
        "rubibceboth": [model.opt_two_bce_both, model.loss_two_bce_both, model.mf_loss_two_bce_both,
                        model.reg_loss_two_bce_both],
        "rubibce": [model.opt_two_bce, model.loss_two_bce, model.mf_loss_two_bce, model.reg_loss_two_bce],
        "normalbce": [model.opt_bce, model.loss_bce, model.mf_loss_bce, model.reg_loss_bce]}[args.train]
    # `--test rubi` ranks with the head of the trained graph (train.py:548-556)
    rubi_type = "rubi_both" if args.train == "rubibceboth" else "rubi_c"
    sess = Session()
    sess.run(global_variables_initializer())
    evaluator = MFEvaluator(data, Ks, args.batch_size, eval_mode=args.eval_mode)

    if args.pretrain != 0:  # train.py:604-612 restores a hard-coded epoch-299 checkpoint
        print("#load existing models.")
        path = os.path.join(ckpt_dir(args), "{}_ckpt.npz".format(299))
        checkpoint.load(path, model)
        model.update_c(sess, args.c)
        users_to_test = list(data.test_user_list.keys())
        ret = evaluator.test(sess, model, users_to_test, model_type="rubi_both")
        print("hit=%.5f recall=%.5f ndcg=%.5f" % (ret["hit_ratio"][0], ret["recall"][0], ret["ndcg"][0]))
        return ret

    config.update(best_hr=0, best_ndcg=0, best_recall=0, best_pre=0, best_epoch=0,
                  best_c_hr=0, best_c_epoch=0, best_c=0.0)
    stopping_step, ret = 0, None
    n_batch = data.n_train // args.batch_size + 1
    stepwise = os.environ.get("MACR_STEPWISE") == "1"
    # The sampler draws from its own RNG stream only, so epoch k+1 can be sampled (native code, GIL
    # released) on a worker thread while epoch k trains and evaluates: same triples in the same
    # order as the serial loop.  The RNG state a checkpoint of epoch k must carry is the one right
    # after epoch k's draws, captured by the worker.
    sampler_pool = None if stepwise else ThreadPoolExecutor(max_workers=1)

    def sample_job():
        b = data.sample_epoch(n_batch)
        return b, (random.getstate(), np.random.get_state())

    pending = None if stepwise else sampler_pool.submit(sample_job)
    rng_after = None
    for epoch in range(args.epoch):
        t1 = time()
        loss, mf_loss, reg_loss = 0.0, 0.0, 0.0
        if stepwise:  # the literal loop of train.py:470-499
            for _ in range(n_batch):
                users, pos_items, neg_items = data.sample()
                _, batch_loss, batch_mf_loss, batch_reg_loss = sess.run(
                    opt_fetches,
                    feed_dict={model.users: users, model.pos_items: pos_items, model.neg_items: neg_items})
                loss += batch_loss / n_batch
                mf_loss += batch_mf_loss / n_batch
                reg_loss += batch_reg_loss / n_batch
        else:
            # same triples (the sampler consumes only its own RNG stream), same steps, same sums:
            # the epoch is sampled natively, staged once and run as one call
            batches, rng_after = pending.result()
            pending = sampler_pool.submit(sample_job) if epoch + 1 < args.epoch else None
            for bl in model.train_epoch(batches):
                loss += float(bl[0]) / n_batch
                mf_loss += float(bl[1]) / n_batch
                reg_loss += float(bl[2]) / n_batch
        if np.isnan(loss):
            print("ERROR: loss is nan.")
            sys.exit()
        if (epoch + 1) % args.log_interval != 0:
            if args.verbose > 0 and epoch % args.verbose == 0:
                perf_str = "Epoch %d [%.1fs]: train==[%.5f=%.5f + %.5f]" % (epoch, time() - t1, loss, mf_loss, reg_loss)
                print(perf_str)
                logging.info(perf_str)
            continue

        t2 = time()
        if args.valid_set == "valid":
            users_to_test = list(data.valid_user_list.keys())
        else:
            users_to_test = list(data.test_user_list.keys())
        if args.test == "rubi" and tune:
            print("Epoch %d" % epoch)
            best = None
            for c in np.linspace(args.start, args.end, args.step):
                model.update_c(sess, c)
                r = evaluator.test(sess, model, users_to_test, model_type=rubi_type, valid_set=args.valid_set)
                t3 = time()
                if args.verbose > 0:
                    perf_str = ("c:%.2f [%.1fs + %.1fs]: train==[%.8f=%.8f + %.8f], recall=[%.5f, %.5f], "
                                "precision=[%.5f, %.5f], hit=[%.5f, %.5f], ndcg=[%.5f, %.5f]") % (
                        c, t2 - t1, t3 - t2, loss, mf_loss, reg_loss, r["recall"][0], r["recall"][-1],
                        r["precision"][0], r["precision"][-1], r["hit_ratio"][0], r["hit_ratio"][-1],
                        r["ndcg"][0], r["ndcg"][-1])
                    print(perf_str)
                    logging.info(perf_str)
                if best is None or r["hit_ratio"][0] > best[1]["hit_ratio"][0]:
                    best = (float(c), r)
            ret = best[1]
            if ret["hit_ratio"][0] >= config["best_hr"]:
                config["best_c"] = best[0]
            head = "best c:%.2f" % best[0]
        elif args.test == "rubi":
            print("Epoch %d" % epoch)
            c = args.c
            model.update_c(sess, c)
            ret = evaluator.test(sess, model, users_to_test, model_type=rubi_type, valid_set=args.valid_set)
            head = "c:%.2f" % c
        elif args.test == "normal":
            ret = evaluator.test(sess, model, users_to_test, model_type="o", valid_set=args.valid_set)
            head = "Epoch %d" % epoch
        else:
            raise SystemExit(f"--test {args.test}: only rubi / normal are implemented")
        t3 = time()
        if args.verbose > 0:
            perf_str = ("%s [%.1fs + %.1fs]: train==[%.8f=%.8f + %.8f], recall=[%.5f, %.5f], "
                        "precision=[%.5f, %.5f], hit=[%.5f, %.5f], ndcg=[%.5f, %.5f]") % (
                head, t2 - t1, t3 - t2, loss, mf_loss, reg_loss, ret["recall"][0], ret["recall"][-1],
                ret["precision"][0], ret["precision"][-1], ret["hit_ratio"][0], ret["hit_ratio"][-1],
                ret["ndcg"][0], ret["ndcg"][-1])
            print(perf_str)
            logging.info(perf_str)
        config, stopping_step, should_stop = early_stop(
            ret["hit_ratio"][0], ret["ndcg"][0], ret["recall"][0], ret["precision"][0], epoch, config,
            stopping_step)
        if args.save_flag == 1:
            checkpoint.save(os.path.join(ckpt_dir(args), "{}_ckpt.npz".format(epoch)), model,
                            {"epoch": epoch}, rng_state=rng_after)
        if should_stop and args.early_stop == 1:
            msg = "{} dataset best epoch{}: hr:{} ndcg:{} recall:{} precision:{}".format(
                args.dataset, config["best_epoch"], config["best_hr"], config["best_ndcg"],
                config["best_recall"], config["best_pre"])
            print(msg)
            logging.info(msg)
            if args.save_flag == 1:
                with open(os.path.join(ckpt_dir(args), "best_epoch.txt"), "w") as f:
                    print(config["best_epoch"], file=f)
                if args.test == "rubi":
                    with open(os.path.join(ckpt_dir(args), "best_c.txt"), "w") as f:
                        print(config["best_c"], file=f)
            break
    if pending is not None:  # early stop: let the speculative draw of the next epoch finish
        pending.result()
    if sampler_pool is not None:
        sampler_pool.shutdown()
    model.close()
    return config


if __name__ == "__main__":
    main()
