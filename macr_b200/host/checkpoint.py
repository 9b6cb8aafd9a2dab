"""Checkpoint / resume: the four parameters, their Adam slots, the Adam step counter and both
host RNG states in one .npz (stand-in for tf.train.Saver, macr_mf/train.py:376,588-602,
macr_lightgcn/LightGCN.py:693-700,891-893).  Unlike the reference, a resumed run continues the
sampler streams and the Adam bias correction exactly where the saved run stopped."""
import os
import pickle
import random

import numpy as np


def save(path, model, extra=None):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    sd = model.state_dict()
    sd["py_random_state"] = np.frombuffer(pickle.dumps(random.getstate()), dtype=np.uint8)
    sd["np_random_state"] = np.frombuffer(pickle.dumps(np.random.get_state()), dtype=np.uint8)
    for k, v in (extra or {}).items():
        sd["extra_" + k] = np.asarray(v)
    tmp = path + ".tmp.npz"
    np.savez(tmp, **sd)
    os.replace(tmp, path if path.endswith(".npz") else path + ".npz")


def load(path, model, restore_rng=True):
    z = np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False)
    model.load_state_dict({k: z[k] for k in z.files if not k.startswith(("py_", "np_", "extra_"))})
    if restore_rng:
        random.setstate(pickle.loads(z["py_random_state"].tobytes()))
        np.random.set_state(pickle.loads(z["np_random_state"].tobytes()))
    return {k[6:]: z[k] for k in z.files if k.startswith("extra_")}
