"""`LightGCN` facade -- the attribute surface of macr_lightgcn/LightGCN.py:32-555 for
`--alg_type lightgcn --loss bceboth --test rubiboth`.

Fetching
  [opt_two_bce_both, loss_two_bce_both, mf_loss_two_bce_both, emb_loss_two_bce_both,
   reg_loss_two_bce_both]  -> one training step: L-layer CSR SpMM propagation, the B x B gated
   BCE on the propagated rows with the L2 term on the raw rows, backward through the layer
   stack, TF-faithful Adam (LightGCN.py:197-201,288-309,495-532); returns
   (None, loss, mf_loss, emb_loss, [0.])                       (reg_loss is a constant, :530)
  [loss_two_bce_both, mf_loss_two_bce_both, emb_loss_two_bce_both] -> the same losses without an
   update (train_thread_test, LightGCN.py:616-647)
  [opt_bce, ...] (`--loss bce`, :183-186,415-429) and [opt_two_bce1, ...] (`--loss bce1`: item gate
   only, :190-194,431-461) step the neighbouring loss graphs with the same kernels
  rubi_ratings_both / rubi_ratings1 / batch_ratings -> score matrix on the propagated tables
   (:166,442,509)
"""
import ast

import numpy as np
import torch

from .. import ops
from .model_mf import _ScoringMixin, init_weights
from .session import Fetch, Placeholder, Unsupported

_TRAIN = ("opt_two_bce_both", "loss_two_bce_both", "mf_loss_two_bce_both", "emb_loss_two_bce_both",
          "reg_loss_two_bce_both")
# `--loss bce` (README.md:59, the baseline): LightGCN.py:183-186,415-429
_TRAIN_BCE = ("opt_bce", "loss_bce", "mf_loss_bce", "emb_loss_bce", "reg_loss_bce")
# `--loss bce1` (item gate only): LightGCN.py:190-194,431-461
_TRAIN_BCE1 = ("opt_two_bce1", "loss_two_bce1", "mf_loss_two_bce1", "emb_loss_two_bce1", "reg_loss_two_bce1")
_UNSUPPORTED = (
    "opt", "loss", "mf_loss", "emb_loss", "reg_loss", "opt_two_bce2", "loss_two_bce2", "mf_loss_two_bce2",
    "emb_loss_two_bce2", "reg_loss_two_bce2", "rubi_ratings2",
    "batch_ratings_causal_c", "direct_minus_ratings_both",
)


class LightGCN(_ScoringMixin):
    def __init__(self, data_config, pretrain_data=None, args=None, device=None):
        if args is None:
            raise ValueError("LightGCN(data_config, pretrain_data, args=...): pass the parsed flags "
                             "(the reference reads a module-global `args`)")
        if args.alg_type != "lightgcn":
            raise NotImplementedError("only --alg_type lightgcn is on the MACR hot path")
        self.model_type, self.adj_type, self.alg_type = "lightgcn", args.adj_type, args.alg_type
        self.n_users, self.n_items = data_config["n_users"], data_config["n_items"]
        self.lr, self.emb_dim, self.batch_size = args.lr, args.embed_size, args.batch_size
        self.weight_size = list(ast.literal_eval(args.layer_size))
        self.n_layers = len(self.weight_size)                      # LightGCN.py:49-50
        self.regs = list(ast.literal_eval(args.regs))
        self.decay = self.regs[0]                                  # :52
        self.alpha, self.beta, self.c = args.alpha, args.beta, args.c
        self.Ks = list(ast.literal_eval(args.Ks))                  # :57
        self.verbose = args.verbose
        self.rubi_c = 0.0
        if self.emb_dim != ops.D:
            raise ops.MacrError(f"--embed_size must be {ops.D} (got {self.emb_dim})")
        dev_index = getattr(args, "device", 0) if device is None else device
        self.dev = torch.device("cuda", dev_index) if isinstance(dev_index, int) else torch.device(dev_index)
        A = data_config["norm_adj"].tocsr().astype(np.float32)     # scipy, :666-681
        A.sort_indices()
        if pretrain_data is not None:                              # --pretrain -1, :222-231
            U = np.asarray(pretrain_data["user_embed"], np.float32)
            I = np.asarray(pretrain_data["item_embed"], np.float32)
            _, _, w, wu = init_weights(self.n_users, self.n_items, self.emb_dim,
                                       getattr(args, "init_seed", 12345))
        else:
            U, I, w, wu = init_weights(self.n_users, self.n_items, self.emb_dim,
                                       getattr(args, "init_seed", 12345), getattr(args, "init_npz", ""))
        if not 0 < int(self.batch_size) <= 8192:
            raise ops.MacrError(f"batch_size {self.batch_size}: the B200 step supports 1..8192")
        self.hp = ops.HParams.make(lr=self.lr, alpha=self.alpha, beta=self.beta, decay=self.decay,
                                   batch_size=self.batch_size)
        self.trainer = ops.LGCNTrainer(A.indptr.astype(np.int32), A.indices.astype(np.int32),
                                       A.data.astype(np.float32), U, I, w, wu, self.n_layers, self.hp,
                                       max_batch=self.batch_size, device=self.dev)
        for name in ("users", "pos_items", "neg_items", "node_dropout", "mess_dropout"):
            setattr(self, name, Placeholder(self, name))
        for name in _TRAIN + _TRAIN_BCE + _TRAIN_BCE1 + ("rubi_ratings_both", "rubi_ratings1", "batch_ratings"):
            setattr(self, name, Fetch(self, name))
        for name in _UNSUPPORTED:
            setattr(self, name, Unsupported(self, name))

    def _run(self, names, feeds):
        if any(n in _TRAIN or n in _TRAIN_BCE or n in _TRAIN_BCE1 for n in names):
            modes = [m for m, group in ((ops.LGCNTrainer.RUBIBCEBOTH, _TRAIN), (ops.LGCNTrainer.NORMALBCE, _TRAIN_BCE),
                                        (ops.LGCNTrainer.RUBIBCE, _TRAIN_BCE1)) if any(n in group for n in names)]
            if len(modes) != 1:
                raise NotImplementedError("fetches of two different loss graphs in one sess.run")
            mode = modes[0]
            if getattr(self, "_mode", ops.LGCNTrainer.RUBIBCEBOTH) != mode:
                self.trainer.set_mode(mode)
                self._mode = mode
            train = any(n.startswith("opt") for n in names)
            loss, mf, emb = self.step(feeds["users"], feeds["pos_items"], feeds["neg_items"], train)
            zero = np.zeros(1, np.float32)  # reg_loss is a constant [0.] (:427, :530)
            val = {"opt_two_bce_both": None, "loss_two_bce_both": loss, "mf_loss_two_bce_both": mf,
                   "emb_loss_two_bce_both": emb, "reg_loss_two_bce_both": zero,
                   "opt_bce": None, "loss_bce": loss, "mf_loss_bce": mf, "emb_loss_bce": emb,
                   "reg_loss_bce": zero, "opt_two_bce1": None, "loss_two_bce1": loss, "mf_loss_two_bce1": mf,
                   "emb_loss_two_bce1": emb, "reg_loss_two_bce1": zero}
            return [val[n] for n in names]
        return [self._run_scores(n, feeds) for n in names]

    def step(self, users, pos_items, neg_items, train=True):
        with torch.cuda.device(self.dev):
            return self.trainer.step_host(users, pos_items, neg_items, train=train)

    _MODE_OF_LOSS = {"bceboth": ops.LGCNTrainer.RUBIBCEBOTH, "bce": ops.LGCNTrainer.NORMALBCE,
                     "bce1": ops.LGCNTrainer.RUBIBCE}

    def run_epoch(self, batches, loss="bceboth", train=True):
        """`batches` int32 [n,3,B] on the host (e.g. `Data.sample_epoch`) -> float32 [n,4] host losses
        {loss, mf_loss, emb_loss, L_ori}: n `sess.run` fetches of the `--loss` graph as ONE call (one
        H2D, n step graphs, one D2H); train=False is the loss-only pass of LightGCN.py:799-819."""
        mode = self._MODE_OF_LOSS[loss]
        if getattr(self, "_mode", ops.LGCNTrainer.RUBIBCEBOTH) != mode:
            self.trainer.set_mode(mode)
            self._mode = mode
        b = torch.as_tensor(np.ascontiguousarray(batches, dtype=np.int32))
        if getattr(self, "_pin", None) is None or self._pin.shape != b.shape:
            self._pin = torch.empty(b.shape, dtype=torch.int32).pin_memory()
        self._pin.copy_(b)
        with torch.cuda.device(self.dev):
            return self.trainer._run_host_train(self._pin, train)

    def _score_tables(self):
        with torch.cuda.device(self.dev):
            ue, ie = self.trainer.embeddings()  # propagated once per parameter version
        t = self.trainer.tab
        return ue, ie, t.w, t.wu

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
