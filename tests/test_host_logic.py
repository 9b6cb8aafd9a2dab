"""Host-side logic that needs no GPU: metric arithmetic against the oracle's restatement of
train.py:32-117, host ranking, the session dispatch rules, checkpoint RNG round trip."""
import random
import types

import numpy as np
import pytest

from macr_b200.host import evaluate, session
from oracle import mf_metrics


def test_mf_metrics_from_hits_matches_reference_arithmetic():
    rng = np.random.RandomState(0)
    T, n_items, Ks = 40, 90, [5, 10, 20]
    topk = np.stack([rng.permutation(n_items)[:20] for _ in range(T)])
    truth = [list(rng.choice(n_items, size=int(rng.randint(1, 30)), replace=False)) for _ in range(T)]
    want = mf_metrics.evaluate(topk, truth, Ks)
    got = evaluate.mf_metrics_from_hits(evaluate._hits(topk, truth), [len(t) for t in truth], Ks)
    for k in want:
        np.testing.assert_allclose(got[k] / T, want[k], rtol=1e-12, atol=1e-15, err_msg=k)


def test_host_topk_masks_and_breaks_ties_by_id():
    rate = np.array([[1.0, 3.0, 3.0, 2.0, 0.5], [5.0, 4.0, 3.0, 2.0, 1.0]], np.float32)
    mrp, mcol = np.array([0, 1, 4], np.int32), np.array([1, 0, 1, 2], np.int32)
    ids = evaluate.host_topk(rate, mrp, mcol, 3)
    assert ids[0].tolist() == [2, 3, 0]
    assert ids[1].tolist() == [3, 4, -1]  # only two unmasked items left


class _FakeModel:
    def __init__(self):
        self.users = session.Placeholder(self, "users")
        self.loss = session.Fetch(self, "loss")
        self.opt = session.Fetch(self, "opt")
        self.other = session.Unsupported(self, "opt_bce")
        self.calls = []

    def _run(self, names, feeds):
        self.calls.append((names, feeds))
        return [None if n == "opt" else 1.5 for n in names]


def test_session_dispatch():
    m, sess = _FakeModel(), session.Session()
    assert sess.run(session.global_variables_initializer()) is None
    out = sess.run([m.opt, m.loss], feed_dict={m.users: [1, 2]})
    assert out == [None, 1.5]
    assert m.calls[-1] == (["opt", "loss"], {"users": [1, 2]})
    assert sess.run(m.loss, {m.users: [3]}) == 1.5
    with pytest.raises(NotImplementedError):
        sess.run([m.other], {m.users: [1]})
    with pytest.raises(ValueError):
        sess.run([m.loss, _FakeModel().loss])


def test_checkpoint_restores_rng_streams(tmp_path):
    from macr_b200.host import checkpoint

    class M:
        def __init__(self):
            self.sd = {"U": np.arange(6, dtype=np.float32).reshape(3, 2), "steps_done": np.int64(7)}

        def state_dict(self):
            return dict(self.sd)

        def load_state_dict(self, sd):
            self.loaded = sd

    random.seed(5)
    np.random.seed(5)
    random.random(), np.random.rand()
    m = M()
    path = str(tmp_path / "ck" / "3_ckpt.npz")
    checkpoint.save(path, m, {"epoch": 3})
    want = (random.random(), np.random.rand())
    random.seed(99)
    np.random.seed(99)
    extra = checkpoint.load(path, m)
    assert (random.random(), np.random.rand()) == want
    assert int(extra["epoch"]) == 3 and int(m.loaded["steps_done"]) == 7
    np.testing.assert_array_equal(m.loaded["U"], m.sd["U"])


def test_drivers_reject_modes_outside_the_hot_path():
    from macr_b200.cli import lightgcn, train_mf

    with pytest.raises(SystemExit):
        train_mf.main(["--dataset", "tiny", "--train", "rubi"])  # BPR two-branch variant: not implemented
    with pytest.raises(SystemExit):
        lightgcn.main(["--dataset", "tiny", "--loss", "bpr"])
