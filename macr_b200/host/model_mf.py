"""`BPRMF` facade -- the attribute surface of macr_mf/model.py:13-326 for the MACR hot path.

What the reference builds as a TF graph (model.py:27-101) is here one resident model on the
GPU: the two embedding tables, the two branch vectors, their Adam slots (all fp32, row-major,
in HBM) behind `ops.MFTrainer`.  Fetching

  [opt_two_bce_both, loss_two_bce_both, mf_loss_two_bce_both, reg_loss_two_bce_both]
        -> one fused training step (gather -> dots -> B x B gated BCE -> row gradients ->
           TF-faithful Adam), returns (None, loss, mf_loss, reg_loss)      model.py:72-74,185-222
  [opt_bce, loss_bce, mf_loss_bce, reg_loss_bce]
        -> one `--train normalbce` step (element-wise BCE, the README's baseline)  model.py:99-101,277-287
  [opt_two_bce, loss_two_bce, mf_loss_two_bce, reg_loss_two_bce]
        -> one `--train rubibce` step (item gate only; w_user untouched)           model.py:83-85,158-183
  rubi_ratings_both  -> ((u.i) - c) * sig(i.w) * sig(u.w_user)  as float32 [B_u, I]  model.py:199
  rubi_ratings       -> ((u.i) - c) * sig(i.w)   (`--train rubibce --test rubi`)      model.py:141
  batch_ratings      -> u.i                                                            model.py:45

Everything else in the reference graph (other --train modes, baselines) is an `Unsupported`
token.  `topk()` is the native evaluation path: fused score + train-item mask + top-K on the
device, nothing but K ids per user comes back.
"""
import numpy as np
import torch

from .. import ops
from .session import Fetch, Placeholder, Unsupported

_TRAIN = ("opt_two_bce_both", "loss_two_bce_both", "mf_loss_two_bce_both", "reg_loss_two_bce_both")
# `--train normalbce` (README.md:30, the baseline the MACR rows are compared with): model.py:99-101
_TRAIN_BCE = ("opt_bce", "loss_bce", "mf_loss_bce", "reg_loss_bce")
# `--train rubibce` (item gate only): model.py:83-85,158-183
_TRAIN_ITEM = ("opt_two_bce", "loss_two_bce", "mf_loss_two_bce", "reg_loss_two_bce")
# score fetch -> head of the fused kernel: which gates multiply (y - c)
_HEAD_OF = {"rubi_ratings_both": "both", "rubi_ratings": "item", "rubi_ratings1": "item", "batch_ratings": "plain"}
_UNSUPPORTED = (
    "opt", "loss", "mf_loss", "reg_loss", "opt_two", "loss_two", "mf_loss_two", "reg_loss_two",
    "opt2", "loss2", "opt2_bce", "loss2_bce", "opt3", "opt3_bce",
    "opt_userc_bce", "loss_userc_bce", "user_const_ratings", "item_const_ratings",
    "user_rand_ratings", "item_rand_ratings", "rubi_ratings_userc",
    "direct_minus_ratings", "direct_minus_ratings_both", "rubi_ratings_both_poptest",
)


def xavier_uniform(rng, rows, cols):
    """tf.contrib.layers.xavier_initializer(): U(-l, l), l = sqrt(6 / (fan_in + fan_out)).
    (TF draws from its own Philox stream; the numbers here come from numpy -- SURVEY 8c.)"""
    lim = np.sqrt(6.0 / (rows + cols))
    return rng.uniform(-lim, lim, size=(rows, cols)).astype(np.float32)


def init_weights(n_users, n_items, d, seed=12345, init_npz=""):
    """user_embedding [U,d], item_embedding [I,d], w [d], w_user [d] (model.py:59-60,107-115)."""
    if init_npz:
        z = np.load(init_npz)
        return (z["U"].astype(np.float32), z["I"].astype(np.float32),
                z["w"].astype(np.float32).reshape(-1), z["wu"].astype(np.float32).reshape(-1))
    rng = np.random.RandomState(seed)
    U = xavier_uniform(rng, n_users, d)
    I = xavier_uniform(rng, n_items, d)
    w = xavier_uniform(rng, d, 1).reshape(-1)
    wu = xavier_uniform(rng, d, 1).reshape(-1)
    return U, I, w, wu


class _ScoringMixin:
    """Full-catalogue counterfactual scoring shared by the MF and LightGCN facades."""

    def _score_tables(self):
        raise NotImplementedError

    def update_c(self, sess, c):
        """model.py:313-314 / LightGCN.py:554-555: rubi_c := c."""
        self.rubi_c = float(c)

    def _ids(self, seq):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(seq, dtype=np.int32))).to(self.dev)

    def _head_c(self, head, c):
        if head not in ("plain", "item", "both"):
            raise ValueError(f"unknown score head {head!r}")
        return 0.0 if head == "plain" else (self.rubi_c if c is None else float(c))

    def _item_gate(self, Iq, w, head):
        if head == "plain":
            return torch.ones(Iq.shape[0], dtype=torch.float32, device=self.dev)
        return ops.score_gates(Iq, w)

    def _user_gate(self, Uq, wu, head):
        if head != "both":
            return torch.ones(Uq.shape[0], dtype=torch.float32, device=self.dev)
        return ops.score_gates(Uq, wu)

    def _gates(self, Uq, Iq, w, wu, head, c):
        """(sig_i, sig_u, c) of a score head: "both" = rubi_ratings_both (model.py:199), "item" =
        rubi_ratings / rubi_ratings1 (model.py:141, LightGCN.py:442; x * 1.0 is exact, so the user
        gate is a vector of ones), "plain" = batch_ratings (((y - 0) * 1) * 1 == y exactly)."""
        cc = self._head_c(head, c)
        return self._item_gate(Iq, w, head), self._user_gate(Uq, wu, head), cc

    def score_matrix(self, users, items=None, c=None, gated=True, head=None):
        """A score head (`head`; `gated` is the older both/plain switch) as a device tensor [B_u, n]."""
        head = head or ("both" if gated else "plain")
        Ut, It, w, wu = self._score_tables()
        with torch.cuda.device(self.dev):
            Uq = ops.gather_rows(Ut, self._ids(users))
            Iq = It if items is None else ops.gather_rows(It, self._ids(items))
            si, su, cc = self._gates(Uq, Iq, w, wu, head, c)
            return ops.score_matrix(Uq, Iq, si, su, cc)

    def topk(self, users, K, mask_rowptr=None, mask_col=None, c=None, head="both", prep=None):
        """Fused score + mask + top-K over the whole catalogue -> (ids [T,K], scores [T,K])
        device tensors; masked = CSR over `users` of item ids to exclude (their train items).
        prep: a dict the caller keeps for as long as the parameters do not change (one evaluation:
        train.py:174-180 / batch_test.py:38-43 walk the test users in batches against one model);
        the item gates and the tensor-core item operands are then prepared by the first batch only."""
        Ut, It, w, wu = self._score_tables()
        with torch.cuda.device(self.dev):
            Uq = ops.gather_rows(Ut, self._ids(users))
            mrp = None if mask_rowptr is None else self._ids(mask_rowptr)
            mcol = None if mask_col is None else self._ids(mask_col)
            if mcol is not None and mcol.numel() == 0:
                mcol = torch.zeros(1, dtype=torch.int32, device=self.dev)
            if prep is None:
                si, su, cc = self._gates(Uq, It, w, wu, head, c)
                return ops.score_topk(Uq, It, si, su, cc, mrp, mcol, K)
            cc = self._head_c(head, c)
            key = (head, cc, It.data_ptr())
            if key not in prep:
                si = self._item_gate(It, w, head)
                items = ops.TcItems(It, si, cc) if ops.uses_tc(It.shape[0], K) else None
                prep[key] = (It, si, items)  # `It` kept alive: the key holds its address
            _, si, items = prep[key]
            su = self._user_gate(Uq, wu, head)
            return ops.score_topk(Uq, It, si, su, cc, mrp, mcol, K, prepared=items)

    def _is_full_range(self, items):
        n = self.n_items
        if isinstance(items, range):
            return items == range(n)
        return len(items) == n and items[0] == 0 and items[-1] == n - 1 and \
            bool(np.array_equal(np.asarray(items), np.arange(n)))

    def _run_scores(self, name, feeds):
        users, items = feeds["users"], feeds["pos_items"]
        sub = None if self._is_full_range(items) else items
        M = self.score_matrix(users, sub, head=_HEAD_OF[name])
        return M.cpu().numpy()


class BPRMF(_ScoringMixin):
    def __init__(self, args, data_config, device=None):
        self.n_users, self.n_items = data_config["n_users"], data_config["n_items"]
        self.decay, self.emb_dim, self.lr = args.regs, args.embed_size, args.lr
        self.batch_size, self.verbose = args.batch_size, args.verbose
        self.c, self.alpha, self.beta = args.c, args.alpha, args.beta
        self.rubi_c = 0.0  # tf.zeros([1]) until update_c (model.py:117)
        if self.emb_dim != ops.D:
            raise ops.MacrError(f"--embed_size must be {ops.D} (got {self.emb_dim})")
        dev_index = getattr(args, "device", 0) if device is None else device
        self.dev = torch.device("cuda", dev_index) if isinstance(dev_index, int) else torch.device(dev_index)
        U, I, w, wu = init_weights(self.n_users, self.n_items, self.emb_dim,
                                   getattr(args, "init_seed", 12345), getattr(args, "init_npz", ""))
        if not 0 < int(self.batch_size) <= 8192:
            raise ops.MacrError(f"batch_size {self.batch_size}: the B200 step supports 1..8192")
        self.hp = ops.HParams.make(lr=self.lr, alpha=self.alpha, beta=self.beta, decay=self.decay,
                                   batch_size=self.batch_size)
        self.trainer = ops.MFTrainer(U, I, w, wu, self.hp, max_batch=self.batch_size,
                                     device=self.dev)
        for name in ("users", "pos_items", "neg_items"):
            setattr(self, name, Placeholder(self, name))
        for name in _TRAIN + _TRAIN_BCE + _TRAIN_ITEM + ("rubi_ratings_both", "rubi_ratings", "batch_ratings"):
            setattr(self, name, Fetch(self, name))
        for name in _UNSUPPORTED:
            setattr(self, name, Unsupported(self, name))
        if self.verbose > 0:  # _statistics_params, model.py:316-326
            total = 2 * (self.n_users + self.n_items) * self.emb_dim + self.emb_dim + self.n_users
            print("#params: %d" % total)

    # ---- session dispatch -----------------------------------------------------------------
    def _run(self, names, feeds):
        if any(n.startswith("opt") for n in names):
            graphs = [g for g, group in (("rubibceboth", _TRAIN), ("normalbce", _TRAIN_BCE), ("rubibce", _TRAIN_ITEM))
                      if any(n in group for n in names)]
            if len(graphs) != 1:
                raise NotImplementedError("one optimizer op per sess.run")
            self.set_train_mode(graphs[0])
            loss, mf, reg = self.train_step(feeds["users"], feeds["pos_items"], feeds["neg_items"])
            val = {"opt_two_bce_both": None, "loss_two_bce_both": loss, "mf_loss_two_bce_both": mf,
                   "reg_loss_two_bce_both": reg, "opt_bce": None, "loss_bce": loss, "mf_loss_bce": mf,
                   "reg_loss_bce": reg, "opt_two_bce": None, "loss_two_bce": loss, "mf_loss_two_bce": mf,
                   "reg_loss_two_bce": reg}
            return [val[n] for n in names]
        if any(n in _TRAIN or n in _TRAIN_BCE or n in _TRAIN_ITEM for n in names):
            raise NotImplementedError("loss fetches without the optimizer op are not used by "
                                      "macr_mf/train.py and are not implemented for MF")
        return [self._run_scores(n, feeds) for n in names]

    def set_train_mode(self, train):
        """which optimizer op steps run: "rubibceboth" (default), "normalbce" or "rubibce"."""
        mode = {"rubibceboth": ops.MFTrainer.RUBIBCEBOTH, "normalbce": ops.MFTrainer.NORMALBCE,
                "rubibce": ops.MFTrainer.RUBIBCE}[train]
        if getattr(self, "_mode", ops.MFTrainer.RUBIBCEBOTH) != mode:
            self.trainer.set_mode(mode)
            self._mode = mode

    def train_step(self, users, pos_items, neg_items):
        """One `rubibceboth` step from host id sequences -> (loss, mf_loss, reg_loss)."""
        with torch.cuda.device(self.dev):
            return self.trainer.step_host(users, pos_items, neg_items)

    def train_epoch(self, batches):
        """`batches` int32 [n,3,B] on the host (e.g. `Data.sample_epoch`) -> float32 [n,4] host
        losses {loss, mf_loss, reg_loss, L_ori}: the inner loop of train.py:470-499 as ONE call
        (one H2D, n step graphs, one D2H) -- same results as n `sess.run` fetches."""
        b = torch.as_tensor(np.ascontiguousarray(batches, dtype=np.int32))
        if getattr(self, "_pin", None) is None or self._pin.shape != b.shape:
            self._pin = torch.empty(b.shape, dtype=torch.int32).pin_memory()
        self._pin.copy_(b)
        with torch.cuda.device(self.dev):
            return self.trainer.run_host(self._pin).numpy()

    def _score_tables(self):
        t = self.trainer.tab
        return t.U, t.I, t.w, t.wu

    # ---- checkpoint (tf.train.Saver stand-in, train.py:376,588-591) ------------------------------
    def state_dict(self):
        sd = self.trainer.tab.state_dict()
        sd["steps_done"] = np.int64(self.trainer.steps_done)
        sd["rubi_c"] = np.float32(self.rubi_c)
        return sd

    def load_state_dict(self, sd):
        t = self.trainer.tab
        with torch.no_grad():
            for k in ("U", "mU", "vU", "I", "mI", "vI", "w", "mw", "vw", "wu", "mwu", "vwu"):
                getattr(t, k).copy_(torch.as_tensor(np.asarray(sd[k], np.float32)).reshape(getattr(t, k).shape))
        self.trainer.set_steps_done(int(sd["steps_done"]))
        self.rubi_c = float(sd.get("rubi_c", 0.0))

    def close(self):
        self.trainer.close()
