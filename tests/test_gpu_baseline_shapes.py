"""-m gpu: trainer-level parity at the BASELINE.json shapes.

* `MFTrainer` vs the oracle at gowalla shape (U=29 858, I=40 981, B=4096 -> the
  `grid_bce_kernel<8,8,2,0,1>` instantiation inside the step graph) and at ml_10m shape
  (U=69 166, I=8 790, B=8192): losses 1e-4, tables 1e-5.
* SURVEY 8d config 1 literally: the real `data/addressa`, triples from the sampler stream
  `random.seed(12345)` (first batch pinned to the reference's own sampler by sha1,
  tests/golden/digests.json), B=1024, alpha=beta=1e-3, regs=1e-5, lr=1e-3; oracle vs GPU after
  1, 10 and 111 (= one epoch) steps.
* `LGCNTrainer` on the real addressa `pre` adjacency (bit-exact vs the reference's get_adj_mat).
* the GPU grid loss against `oracle/literal_torch.bce_two_branch_both` DIRECTLY at B=4096
  (the literal [B,B] broadcast graph, macr_mf/model.py:185-222), not via the C oracle.
Reference: macr_mf/model.py:185-222, macr_mf/train.py:464-499.
"""
import hashlib
import json
import os
import random
import types

import numpy as np
import pytest

from helpers import make_batch, make_model

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADDRESSA = os.path.join(ROOT, "data", "addressa")


@pytest.fixture(scope="module")
def T():
    import torch

    return torch


@pytest.fixture(scope="module")
def ops():
    from macr_b200 import ops as o

    return o


def dev(T, a):
    return T.as_tensor(np.ascontiguousarray(a)).cuda()


def _check_tables(tr, st, atol=1e-5):
    t = tr.tab
    for name in ("U", "I", "w", "wu", "mU", "vU", "mI", "vI"):
        np.testing.assert_allclose(getattr(t, name).cpu().numpy(), getattr(st, name), rtol=1e-4, atol=atol,
                                   err_msg=name)


@pytest.mark.parametrize("name,n_users,n_items,B,alpha", [("gowalla", 29858, 40981, 4096, 1e-2),
                                                          ("ml_10m", 69166, 8790, 8192, 1e-3)])
def test_mf_trainer_at_baseline_shapes(T, ops, oracle, name, n_users, n_items, B, alpha):
    steps = 3
    U, I, w, wu = make_model(len(name), n_users, n_items, scale=3.0)
    hp = dict(lr=1e-3, alpha=alpha, beta=1e-3, decay=1e-5, batch_size=B)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.MFTrainer(U, I, w, wu, ops.HParams.make(**hp), max_batch=B)
    rng = np.random.RandomState(B)
    batches = np.stack([np.stack(make_batch(rng, n_users, n_items, B)) for _ in range(steps)]).astype(np.int32)
    want = np.stack([oracle.mf_step(st, *batches[s], oracle.HParams.make(**hp)) for s in range(steps)])
    got = tr.run(dev(T, batches)).cpu().numpy()  # the captured step graph, epoch replay
    np.testing.assert_allclose(got[:, :3], want[:, :3], rtol=1e-4, atol=1e-4)
    _check_tables(tr, st)
    assert np.abs(st.U - U).max() > 0 and tr.steps_done == steps
    tr.close()


def _addressa_args():
    return types.SimpleNamespace(dataset="addressa", data_path=os.path.join(ROOT, "data") + "/", batch_size=1024,
                                 valid_set="test", data_type="ori", source="normal", model="mf")


@pytest.mark.skipif(not os.path.exists(os.path.join(ADDRESSA, "train.txt")), reason="data/addressa not staged")
def test_config1_addressa_reference_sampler_triples_1_10_111_steps(T, ops, oracle, monkeypatch):
    from macr_b200.host.data_mf import Data

    monkeypatch.chdir(ROOT)  # the loader reads ./data/<dataset>/ like the reference (load_data.py:27)
    data = Data(_addressa_args())
    n_batch = data.n_train // 1024 + 1
    assert (data.n_users, data.n_items, n_batch) == (13485, 744, 111)
    random.seed(12345)
    epoch = data.sample_epoch(n_batch)  # int32 [111,3,1024], the reference's sampler stream
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "digests.json")))["mf_addressa"]
    first = hashlib.sha1(np.array([epoch[0, 0], epoch[0, 1], epoch[0, 2]], np.int64).tobytes()).hexdigest()
    assert first == want["first_batch_sha1"]
    # tables from torch.Generator().manual_seed(12345) uniform +-sqrt(6/(rows+64))
    g = T.Generator().manual_seed(12345)
    uni = lambda rows, cols, lim: ((T.rand((rows, cols), generator=g) * 2 - 1) * lim).numpy().astype(np.float32)
    U = uni(data.n_users, 64, np.sqrt(6.0 / (data.n_users + 64)))
    I = uni(data.n_items, 64, np.sqrt(6.0 / (data.n_items + 64)))
    w, wu = uni(64, 1, np.sqrt(6.0 / 65)).ravel(), uni(64, 1, np.sqrt(6.0 / 65)).ravel()
    hp = dict(lr=1e-3, alpha=1e-3, beta=1e-3, decay=1e-5, batch_size=1024)
    st = oracle.MFState(U, I, w, wu)
    tr = ops.MFTrainer(U, I, w, wu, ops.HParams.make(**hp), max_batch=1024)
    hp_o = oracle.HParams.make(**hp)
    done = 0
    for upto in (1, 10, 111):
        want_l = np.stack([oracle.mf_step(st, *epoch[s], hp_o) for s in range(done, upto)])
        got_l = tr.run_host(T.from_numpy(epoch[done:upto].copy())).numpy()  # the CLI's epoch call
        np.testing.assert_allclose(got_l[:, :3], want_l[:, :3], rtol=1e-4, atol=1e-4, err_msg=f"steps {done}..{upto}")
        _check_tables(tr, st)
        done = upto
    tr.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(ADDRESSA, "train.txt")), reason="data/addressa not staged")
def test_lgcn_trainer_on_the_real_addressa_adjacency(T, ops, oracle):
    import contextlib
    import io

    from macr_b200.host.data_lgcn import Data

    B, L, steps = 1024, 2, 4
    with contextlib.redirect_stdout(io.StringIO()):
        data = Data(ADDRESSA, B)
        rowptr, col, val = data.adj_csr("pre")
    assert len(col) == 226690 and data.n_users + data.n_items == 14229
    random.seed(12345)
    np.random.seed(12345)
    U, I, w, wu = make_model(3, data.n_users, data.n_items, scale=3.0)
    hp = dict(lr=1e-3, alpha=1e-2, beta=1e-3, decay=1e-4, batch_size=B)  # README.md:82
    st = oracle.MFState(U, I, w, wu)
    tr = ops.LGCNTrainer(rowptr, col, val, U, I, w, wu, L, ops.HParams.make(**hp), max_batch=B)
    hp_o = oracle.HParams.make(**hp)
    for s in range(steps):
        u, p, n = (np.asarray(x, np.int32) for x in data.sample())
        lo = oracle.lgcn_step(st, rowptr, col, val, L, u, p, n, hp_o, train=True)
        lg = np.array(tr.step_host(u.tolist(), p.tolist(), n.tolist()))
        np.testing.assert_allclose(lg, lo[:3], rtol=1e-4, atol=1e-4, err_msg=f"step {s}")
    t = tr.tab
    for name in ("U", "I", "w", "wu"):
        np.testing.assert_allclose(getattr(t, name).cpu().numpy(), getattr(st, name), rtol=1e-4, atol=1e-5,
                                   err_msg=name)
    ue, ie = tr.embeddings()
    np.testing.assert_allclose(T.cat([ue, ie]).cpu().numpy(),
                               oracle.lgcn_propagate(rowptr, col, val, st.U, st.I, L), rtol=1e-4, atol=1e-5)
    tr.close()


def test_grid_loss_against_the_literal_graph_at_B4096(T, ops):
    """GPU gather + B x B grid vs the literal torch restatement of model.py:185-222 (fp64 on the CPU:
    the [4096,4096] broadcast tensors exist for real), loss tolerance 1e-4 (north_star)."""
    from oracle import literal_torch

    n_users, n_items, B = 29858, 40981, 4096
    U, I, w, wu = make_model(21, n_users, n_items, scale=4.0)
    u, p, n = make_batch(np.random.RandomState(4), n_users, n_items, B)
    alpha, beta, decay = 1e-2, 1e-3, 1e-5
    t64 = lambda a: T.from_numpy(np.asarray(a, np.float64))
    mf, reg, l_ori, l_item, l_user = literal_torch.bce_two_branch_both(
        t64(U[u]), t64(I[p]), t64(I[n]), t64(w).reshape(-1, 1), t64(wu).reshape(-1, 1), alpha, beta, decay, B)
    dU, dI = dev(T, U), dev(T, I)
    yp, yn, sp, sn, su, regsq = ops.gather_dots(dU, dI, dU, dI, dev(T, w), dev(T, wu), dev(T, u), dev(T, p),
                                                dev(T, n))
    l3, _ = ops.grid_bce(yp, yn, sp, sn, su, alpha, beta)
    l3 = l3.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(l3, [float(l_ori), float(l_item), float(l_user)], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(l3[0] + alpha * l3[1] + beta * l3[2], float(mf), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(decay * 0.5 * float(regsq.double().sum().item()) / B, float(reg), rtol=1e-5)
    # and the whole step's reported (loss, mf, reg) against the same literal numbers
    tr = ops.MFTrainer(U, I, w, wu, ops.HParams.make(lr=1e-3, alpha=alpha, beta=beta, decay=decay, batch_size=B),
                       max_batch=B)
    loss, mf_g, reg_g = tr.step_host(u.tolist(), p.tolist(), n.tolist())
    np.testing.assert_allclose([loss, mf_g, reg_g], [float(mf + reg), float(mf), float(reg)], rtol=1e-4, atol=1e-4)
    tr.close()
