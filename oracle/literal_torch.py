"""Literal torch restatement of the reference's TF-1.14 graph (autograd does what TF autodiff
does).  TEST INFRASTRUCTURE ONLY -- used to pin the closed-form gradients of macr_oracle.c.

Every line cites the reference statement it restates; tensor shapes are kept exactly,
including the ``[B] * [B,1] -> [B,B]`` broadcast of macr_mf/model.py:204-205.
"""
import torch


def bce_two_branch_both(users, pos_items, neg_items, w, w_user, alpha, beta, decay,
                        batch_size, reg_rows=None):
    """macr_mf/model.py:185-222 (create_bce_loss_two_brach_both).

    users/pos_items/neg_items: gathered rows [B,d]; w, w_user: [d,1].
    reg_rows: (u,p,n) raw rows for the L2 term (LightGCN.py:525-526); default = same rows.
    Returns (mf_loss, reg_loss, mf_loss_ori, mf_loss_item, mf_loss_user).
    """
    pos_scores = torch.sum(users * pos_items, dim=1)                      # :186  [B]
    neg_scores = torch.sum(users * neg_items, dim=1)                      # :187  [B]
    pos_item_scores = pos_items @ w                                       # :194  [B,1]
    neg_item_scores = neg_items @ w                                       # :195  [B,1]
    user_scores = users @ w_user                                          # :196  [B,1]
    pos_scores = pos_scores * torch.sigmoid(pos_item_scores) * torch.sigmoid(user_scores)  # :204 [B,B]
    neg_scores = neg_scores * torch.sigmoid(neg_item_scores) * torch.sigmoid(user_scores)  # :205 [B,B]
    mf_loss_ori = torch.mean(-torch.log(torch.sigmoid(pos_scores) + 1e-10)
                             - torch.log(1 - torch.sigmoid(neg_scores) + 1e-10))          # :211
    mf_loss_item = torch.mean(-torch.log(torch.sigmoid(pos_item_scores) + 1e-10)
                              - torch.log(1 - torch.sigmoid(neg_item_scores) + 1e-10))    # :213
    mf_loss_user = torch.mean(-torch.log(torch.sigmoid(user_scores) + 1e-10)
                              - torch.log(1 - torch.sigmoid(user_scores) + 1e-10))        # :215
    mf_loss = mf_loss_ori + alpha * mf_loss_item + beta * mf_loss_user                     # :217
    ru, rp, rn = reg_rows if reg_rows is not None else (users, pos_items, neg_items)
    l2 = lambda x: torch.sum(x * x) / 2                                   # tf.nn.l2_loss
    regularizer = (l2(ru) + l2(rp) + l2(rn)) / batch_size                 # :219-220
    reg_loss = decay * regularizer                                        # :221
    return mf_loss, reg_loss, mf_loss_ori, mf_loss_item, mf_loss_user


def bce_plain(users, pos_items, neg_items, decay, batch_size):
    """macr_mf/model.py:277-287 (create_bce_loss, `--train normalbce`)."""
    pos_scores = torch.sum(users * pos_items, dim=1)                      # :278
    neg_scores = torch.sum(users * neg_items, dim=1)                      # :279
    mf_loss = torch.mean(-torch.log(torch.sigmoid(pos_scores) + 1e-9)
                         - torch.log(1 - torch.sigmoid(neg_scores) + 1e-9))               # :282
    l2 = lambda x: torch.sum(x * x) / 2
    regularizer = (l2(users) + l2(pos_items) + l2(neg_items)) / batch_size                # :284-285
    return mf_loss, decay * regularizer                                                   # :286


def bce_two_branch(users, pos_items, neg_items, w, alpha, decay, batch_size, reg_rows=None):
    """macr_mf/model.py:158-183 (create_bce_loss_two_brach, `--train rubibce`; LightGCN.py:431-461
    `--loss bce1` is the same with the L2 term on the raw rows)."""
    pos_scores = torch.sum(users * pos_items, dim=1)                      # :159
    neg_scores = torch.sum(users * neg_items, dim=1)                      # :160
    pos_item_scores = pos_items @ w                                       # :166  [B,1]
    neg_item_scores = neg_items @ w                                       # :167
    pos_scores = pos_scores * torch.sigmoid(pos_item_scores)              # :172  [B]*[B,1] -> [B,B]
    neg_scores = neg_scores * torch.sigmoid(neg_item_scores)              # :173
    mf_loss_ori = torch.mean(-torch.log(torch.sigmoid(pos_scores) + 1e-10)
                             - torch.log(1 - torch.sigmoid(neg_scores) + 1e-10))          # :174
    mf_loss_item = torch.mean(-torch.log(torch.sigmoid(pos_item_scores) + 1e-10)
                              - torch.log(1 - torch.sigmoid(neg_item_scores) + 1e-10))    # :176
    mf_loss = mf_loss_ori + alpha * mf_loss_item                                          # :178
    l2 = lambda x: torch.sum(x * x) / 2
    ru, rp, rn = reg_rows if reg_rows is not None else (users, pos_items, neg_items)      # LightGCN: raw rows
    regularizer = (l2(ru) + l2(rp) + l2(rn)) / batch_size                                 # :180-181
    return mf_loss, decay * regularizer, mf_loss_ori, mf_loss_item


def rubi_ratings(user_rows, item_rows, w, c):
    """macr_mf/model.py:45,141: (batch_ratings - rubi_c) * squeeze(sigmoid(items@w))  (`rubi_c` head)."""
    batch_ratings = user_rows @ item_rows.t()
    return (batch_ratings - c) * torch.sigmoid(item_rows @ w).squeeze(-1)


def rubi_ratings_both(user_rows, item_rows, w, w_user, c):
    """macr_mf/model.py:45,199: (batch_ratings - rubi_c) * sigmoid(items@w)^T * sigmoid(users@w_user)."""
    batch_ratings = user_rows @ item_rows.t()
    return (batch_ratings - c) * torch.sigmoid(item_rows @ w).t() * torch.sigmoid(user_rows @ w_user)


def lightgcn_embed(A, U, I, n_layers):
    """macr_lightgcn/LightGCN.py:288-309 with A a torch sparse/dense [N,N] matrix."""
    ego = torch.cat([U, I], 0)
    allE = [ego]
    for _ in range(n_layers):
        ego = A @ ego
        allE.append(ego)
    mean = torch.mean(torch.stack(allE, 1), dim=1)
    return mean[: U.shape[0]], mean[U.shape[0]:]


class TFAdam:
    """TF-1.14 Adam on torch tensors. sparse=True -> adam.py _apply_sparse_shared after
    _deduplicate_indexed_slices; sparse=False -> training_ops ApplyAdam."""

    def __init__(self, params, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.params, self.lr, self.b1, self.b2, self.eps = params, lr, beta1, beta2, eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.b1p, self.b2p = beta1, beta2

    def step(self, grads, sparse_flags):
        lr_t = self.lr * (1 - self.b2p) ** 0.5 / (1 - self.b1p)
        with torch.no_grad():
            for p, g, m, v, sp in zip(self.params, grads, self.m, self.v, sparse_flags):
                if sp:
                    m.mul_(self.b1).add_(g * (1 - self.b1))
                    v.mul_(self.b2).add_(g * g * (1 - self.b2))
                else:
                    m.add_((g - m) * (1 - self.b1))
                    v.add_((g * g - v) * (1 - self.b2))
                p.sub_(lr_t * m / (v.sqrt() + self.eps))
        self.b1p *= self.b1
        self.b2p *= self.b2
