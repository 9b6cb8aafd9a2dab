import sys, time, random, types, contextlib, io
sys.path.insert(0, '/root/repo')
import numpy as np
from macr_b200.host.data_mf import Data
for ds, B in (("gowalla", 4096), ("ml_10m", 8192), ("addressa", 1024)):
    args = types.SimpleNamespace(data_path="/root/reference/data/", dataset=ds, batch_size=B, data_type="ori",
                                 model="mf", source="normal", valid_set="test")
    with contextlib.redirect_stdout(io.StringIO()):
        data = Data(args)
    random.seed(1)
    data.sample_epoch(2)
    n = 100
    t0 = time.perf_counter(); out = data.sample_epoch(n); t1 = time.perf_counter()
    print("%s B=%d: %.3f ms/batch  %.1f M triples/s  checksum %d" % (ds, B, (t1 - t0) / n * 1e3, B * n / (t1 - t0) / 1e6,
                                                                   int(out.astype(np.int64).sum())))
