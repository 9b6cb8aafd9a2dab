"""Host-side timing of the epoch samplers (no GPU work): best of several epochs per data set,
per-batch twins beside the epoch forms.  python dev/sampler_bench.py"""
import contextlib, io, os, random, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from macr_b200.host.data_mf import Data
from macr_b200.host import native_sampler as ns
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for ds, B in (("gowalla", 4096), ("ml_10m", 8192), ("addressa", 1024)):
    path = os.path.join(ROOT, "data") if os.path.isdir(os.path.join(ROOT, "data", ds)) else "/root/reference/data"
    if not os.path.isdir(os.path.join(path, ds)):
        continue
    args = types.SimpleNamespace(data_path=path + "/", dataset=ds, batch_size=B, data_type="ori",
                                 model="mf", source="normal", valid_set="test")
    cwd = os.getcwd()
    os.chdir(ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        data = Data(args)
    os.chdir(cwd)
    n = data.n_train // B + 1
    random.seed(1)
    data.sample_epoch(2)
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter(); out = data.sample_epoch(n); best = min(best, time.perf_counter() - t0)
    random.seed(1)
    data.sample_epoch(2)
    t0 = time.perf_counter()
    for k in range(min(n, 40)):
        ns.sample_mf(data._ns_pop, data.n_users, data.n_items, data._ns_csr, B)
    per_batch = (time.perf_counter() - t0) / min(n, 40)
    print("%s B=%d: epoch form %.2f ms per epoch of %d batches = %.1f ns/triple (%.1f M triples/s); per-batch twin %.1f ns/triple; "
          "checksum %d" % (ds, B, best * 1e3, n, best / n / B * 1e9, B * n / best / 1e6, per_batch / B * 1e9,
                           int(out.astype(np.int64).sum())))
