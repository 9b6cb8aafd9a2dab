"""Command-line flags of the two reference drivers, same names / types / defaults.

  MF        macr_mf/parse.py:3-92                 (`python train.py --dataset ... --train rubibceboth`)
  LightGCN  macr_lightgcn/utility/parser.py:10-104 (`python LightGCN.py --loss bceboth ...`)

Table-driven: one row per flag (name, type-or-None for free strings, default).  List-valued
flags stay strings on the namespace exactly like the reference (it `eval()`s them at the use
site); use ``as_list`` -- ``ast.literal_eval`` -- to read them.
"""
import argparse
import ast

# (flag, type, default); type None = string flag declared with nargs='?'
MF_FLAGS = [
    ("data_path", None, "./data/"), ("dataset", None, "movielens_ml_1m"), ("source", None, "normal"),
    ("train", None, "normalbce"), ("test", None, "normal"), ("valid_set", None, "test"),
    ("alpha", float, 1e-3), ("beta", float, 1e-3), ("early_stop", int, 1), ("verbose", int, 1),
    ("epoch", int, 1000), ("embed_size", int, 64), ("batch_size", int, 1024), ("Ks", None, "[20]"),
    ("epochs", None, "[]"), ("regs", float, 1e-5), ("c", float, 40.0), ("train_c", str, "val"),
    ("lr", float, 1e-3), ("wd", float, 1e-5), ("model", None, "mf"), ("skew", int, 0),
    ("devide_ratio", float, 0.8), ("save_flag", int, 1), ("cuda", str, "1"), ("pretrain", int, 0),
    ("check_c", int, 1), ("log_interval", int, 10), ("pop_wd", float, 0.0), ("base", float, -1.0),
    ("cf_pen", float, 1.0), ("saveID", None, ""), ("user_min", int, 1), ("user_max", int, 1000),
    ("data_type", None, "ori"), ("imb_type", None, "exp"), ("top_ratio", float, 0.1),
    ("lam", float, 1.0), ("check_epoch", None, "all"), ("start", float, -1.0), ("end", float, 1.0),
    ("step", int, 20), ("out", int, 0),
]

LGCN_FLAGS = [
    ("weights_path", None, ""), ("data_path", None, "../data/"), ("proj_path", None, ""),
    ("dataset", None, "gowalla"), ("valid_set", None, "test"), ("pretrain", int, 0),
    ("verbose", int, 1), ("is_norm", int, 1), ("epoch", int, 1000), ("embed_size", int, 64),
    ("layer_size", None, "[64, 64, 64, 64]"), ("batch_size", int, 1024),
    ("regs", None, "[1e-5,1e-5,1e-2]"), ("lr", float, 0.01), ("c", float, 40.0),
    ("model_type", None, "lightgcn"), ("adj_type", None, "pre"), ("alg_type", None, "lightgcn"),
    ("gpu_id", int, 0), ("node_dropout_flag", int, 0), ("node_dropout", None, "[0.1]"),
    ("mess_dropout", None, "[0.1]"), ("Ks", None, "[1,5,10,15,20,30]"), ("save_flag", int, 1),
    ("test_flag", None, "part"), ("saveID", None, ""), ("base", float, -1.0),
    ("log_interval", int, 10), ("only_test", int, 0), ("loss", None, "bpr"), ("alpha", float, 1e-3),
    ("beta", float, 1e-3), ("test", None, "normal"), ("early_stop", int, 1), ("start", float, -1.0),
    ("end", float, 1.0), ("step", int, 20), ("out", int, 0),
]

# flags this build adds (none of them changes results; all default to "off")
EXTRA_FLAGS = [
    ("device", int, 0),          # CUDA device index (the reference uses CUDA_VISIBLE_DEVICES)
    ("init_seed", int, 12345),   # Xavier draw (TF's Philox stream is not reproducible)
    ("init_npz", None, ""),      # load U/I/w/w_user from an .npz instead of drawing them
    ("eval_mode", None, "fused"),  # fused = on-device score+mask+top-K; matrix = literal fetch
]


def _build(description, table):
    parser = argparse.ArgumentParser(description=description)
    for name, typ, default in table + EXTRA_FLAGS:
        if typ is None:
            parser.add_argument("--" + name, nargs="?", default=default)
        else:
            parser.add_argument("--" + name, type=typ, default=default)
    return parser


def mf_parser():
    return _build("Run pop_bias.", MF_FLAGS)


def lgcn_parser():
    return _build("Run NGCF.", LGCN_FLAGS)


def parse_mf_args(argv=None):
    return mf_parser().parse_args(argv)


def parse_lgcn_args(argv=None):
    return lgcn_parser().parse_args(argv)


def as_list(text):
    """`eval(args.Ks)` of the reference, without eval."""
    value = ast.literal_eval(text) if isinstance(text, str) else text
    return list(value) if isinstance(value, (list, tuple)) else [value]


MAX_BATCH, MAX_TOPK = 8192, 32  # macr_*_trainer_create (csrc/trainer.cu), kMaxKFast / the tcgen05 re-rank (csrc/score*.cu)


def check_device_limits(batch_size, Ks):
    """Fail at start-up, not at the first step / first evaluation, when a flag exceeds what the
    device path was built for."""
    if not 0 < int(batch_size) <= MAX_BATCH:
        raise SystemExit(f"--batch_size {batch_size}: the B200 step supports 1..{MAX_BATCH}")
    if max(Ks) > MAX_TOPK or min(Ks) < 1:
        raise SystemExit(f"--Ks {list(Ks)}: the fused top-K supports cut-offs in 1..{MAX_TOPK}")
