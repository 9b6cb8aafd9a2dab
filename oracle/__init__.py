"""CPU oracle for MACR's hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``macr_b200``) never does:
its ops fail loudly when the CUDA library is missing.

Parity status: **parity unpinned** at the TensorFlow-1.14 boundary (the reference has no
tests / golden vectors and TF 1.14 cannot run here); see ``macr_oracle.h`` for what *is*
pinned (autograd of the literal restatement, the reference's own C++ evaluator built into
``oracle/_ref``, the reference's own samplers / adjacency via ``tests/golden``).
"""
from .capi import (  # noqa: F401
    HParams,
    MFState,
    lib,
    build,
    gather_dots,
    grid_bce,
    adam_sparse,
    adam_dense_vec,
    adam_lr_t,
    mf_step,
    mf_step_normal,
    mf_step_item,
    grid_bce_item,
    spmm_csr,
    lgcn_propagate,
    lgcn_step,
    lgcn_step_normal,
    lgcn_step_item,
    score_gates,
    score_matrix,
    score_topk,
    topk_rows,
    topk_merge,
    foldout_metrics,
    inv_log2_table,
    set_threads,
    get_threads,
)
