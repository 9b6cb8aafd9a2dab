"""Session-style fetch surface of the reference drivers (TF-1.14 `sess.run(fetches, feed_dict)`).

The reference's train / test loops address the model through *attributes that are graph
tensors*: `sess.run([model.opt_two_bce_both, model.loss_two_bce_both, ...], {model.users: ...})`
(macr_mf/train.py:492-496, macr_lightgcn/LightGCN.py:598-607) and
`sess.run(model.rubi_ratings_both, {model.users: batch, model.pos_items: range(ITEM_NUM)})`
(train.py:249-251, utility/batch_test.py:85-88).  Here those attributes are `Fetch` /
`Placeholder` tokens and `Session.run` hands the request to the owning model, which executes
it on the GPU through libmacr_b200.so -- so a driver written against the reference's surface
runs unchanged.  Callable from any host thread (LightGCN.py:582-614 trains from a worker
thread): ctypes releases the GIL for the duration of every library call.
"""


class Placeholder:
    """`tf.placeholder` stand-in: a key of the feed dict."""

    def __init__(self, owner, name):
        self.owner, self.name = owner, name

    def __repr__(self):
        return f"<placeholder {self.name}>"


class Fetch:
    """A fetchable graph node of the reference model (`model.<name>`)."""

    def __init__(self, owner, name):
        self.owner, self.name = owner, name

    def __repr__(self):
        return f"<fetch {self.name}>"


class Unsupported(Fetch):
    """A node of the reference graph that is outside the MACR hot path (other --train / --loss
    modes, baselines).  Fetching it fails loudly instead of silently computing something else."""


class Session:
    """`tf.Session` stand-in: `run(fetches, feed_dict)` -> values in fetch order."""

    def __init__(self, config=None):
        self.config = config

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        if all(f is None or isinstance(f, _NoOp) for f in flist):
            return None if single else [None] * len(flist)
        owners = {id(f.owner): f.owner for f in flist if isinstance(f, Fetch)}
        if len(owners) != 1:
            raise ValueError("Session.run: fetches must belong to exactly one model")
        for f in flist:
            if isinstance(f, Unsupported):
                raise NotImplementedError(
                    f"{f.name}: only the MACR hot path is implemented (MF --train rubibceboth / "
                    f"LightGCN --loss bceboth and their --test rubi scoring heads); see DESIGN.md")
        owner = next(iter(owners.values()))
        feeds = {}
        for key, value in (feed_dict or {}).items():
            feeds[key.name if isinstance(key, Placeholder) else key] = value
        out = owner._run([f.name for f in flist], feeds)
        return out[0] if single else out

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _NoOp:
    """`tf.global_variables_initializer()` stand-in (tables are initialised at construction)."""


def global_variables_initializer():
    return _NoOp()
