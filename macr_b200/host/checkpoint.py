"""Checkpoint / resume: the four parameters, their Adam slots, the Adam step counter and both
host RNG states in one .npz (stand-in for tf.train.Saver, macr_mf/train.py:376,588-602,
macr_lightgcn/LightGCN.py:693-700,891-893).  Unlike the reference, a resumed run continues the
sampler streams and the Adam bias correction exactly where the saved run stopped."""
import os
import random

import numpy as np


def _rng_arrays(rng_state=None):
    """Both host RNG states as plain arrays (no pickle in either direction): CPython's
    `random` state is (version, 625 ints = MT19937 key + position, gauss_next); numpy's legacy
    state is ('MT19937', uint32[624], pos, has_gauss, cached_gaussian).  rng_state: an explicit
    (random.getstate(), np.random.get_state()) pair instead of the interpreter's current one (a
    sampler that runs ahead on a worker thread hands in the states as of the epoch being saved)."""
    ver, key, gauss = rng_state[0] if rng_state else random.getstate()
    name, npkey, pos, has_gauss, cached = rng_state[1] if rng_state else np.random.get_state()
    if name != "MT19937":
        raise ValueError(f"unsupported numpy bit generator {name!r}")
    return {"py_random_version": np.asarray(ver, np.int64),
            "py_random_key": np.asarray(key, np.uint32),
            "py_random_gauss": np.asarray([0.0 if gauss is None else gauss, gauss is not None], np.float64),
            "np_random_key": np.asarray(npkey, np.uint32),
            "np_random_pos": np.asarray([pos, has_gauss], np.int64),
            "np_random_cached": np.asarray(cached, np.float64)}


def _restore_rng(z):
    if "py_random_key" not in z.files:
        raise ValueError("checkpoint carries no (or a pickled, unsupported) RNG state; "
                         "load it with restore_rng=False")
    g = z["py_random_gauss"]
    random.setstate((int(z["py_random_version"]), tuple(int(x) for x in z["py_random_key"]),
                     float(g[0]) if g[1] else None))
    p = z["np_random_pos"]
    np.random.set_state(("MT19937", z["np_random_key"].astype(np.uint32), int(p[0]), int(p[1]),
                         float(z["np_random_cached"])))


def save(path, model, extra=None, rng_state=None):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    sd = model.state_dict()
    sd.update(_rng_arrays(rng_state))
    for k, v in (extra or {}).items():
        sd["extra_" + k] = np.asarray(v)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **sd)
    os.replace(tmp, path if path.endswith(".npz") else path + ".npz")


def load(path, model, restore_rng=True):
    z = np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False)
    model.load_state_dict({k: z[k] for k in z.files if not k.startswith(("py_", "np_", "extra_"))})
    if restore_rng:
        _restore_rng(z)
    return {k[6:]: z[k] for k in z.files if k.startswith("extra_")}
