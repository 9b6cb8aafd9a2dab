"""ctypes binding of libmacr_b200.so -- the C ABI declared in include/macr_b200.h.

There is NO fallback: if the shared library is missing or a call fails, a ``MacrError`` is
raised.  Device memory / streams come from torch (plumbing only); every pointer handed to the
library is a raw ``tensor.data_ptr()``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MACR_B200_LIB: developer override (A/B of two builds of the same ABI); never a fallback
LIB_PATH = os.environ.get("MACR_B200_LIB") or os.path.join(_HERE, "libmacr_b200.so")
_LIB = None

vp = C.c_void_p
i64 = C.c_int64
i32 = C.c_int
f32 = C.c_float
sz = C.c_size_t


class ShardDesc(C.Structure):
    """macr_shard_desc of include/macr_b200.h."""

    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("u_lo", i64), ("u_hi", i64),
                ("i_lo", i64), ("i_hi", i64), ("max_batch", C.c_int32)]


class MacrError(RuntimeError):
    pass


class HParams(C.Structure):
    """macr_hparams of include/macr_b200.h."""

    _fields_ = [("lr", f32), ("beta1", f32), ("beta2", f32), ("eps", f32), ("alpha", f32),
                ("beta", f32), ("decay", f32), ("batch_size_flag", C.c_int32)]

    @classmethod
    def make(cls, lr=1e-3, alpha=1e-3, beta=1e-3, decay=1e-5, batch_size=1024, beta1=0.9,
             beta2=0.999, eps=1e-8):
        return cls(lr, beta1, beta2, eps, alpha, beta, decay, batch_size)


# name -> (restype, argtypes); must list every symbol include/macr_b200.h declares
PROTOTYPES = {
    "macr_last_error": (C.c_char_p, []),
    "macr_abi_version": (i32, []),
    "macr_device_sm_count": (i32, [C.POINTER(i32)]),
    "macr_gather_dots": (i32, [vp] * 9 + [i32, i32] + [vp] * 6 + [vp]),
    "macr_grid_bce_workspace_bytes": (sz, [i32]),
    "macr_grid_bce_fwd_bwd": (i32, [vp] * 5 + [i32, f32, f32] + [vp] * 6 + [vp, sz, vp]),
    "macr_batch_plan_workspace_bytes": (sz, [i32]),
    "macr_batch_plan": (i32, [vp, i32, i64] + [vp] * 5 + [vp, sz, vp]),
    "macr_adam_sweep_untouched": (i32, [vp, vp, vp, i64, i32, vp, f32, f32, f32, f32, vp]),
    "macr_adam_rows": (i32, [vp, vp, vp, i64, i32, vp, vp, i32, vp, f32, f32, f32, f32, vp]),
    "macr_adam_dense": (i32, [vp, vp, vp, vp, i64, i32, f32, f32, f32, f32, vp]),
    "macr_adam_vec": (i32, [vp, vp, vp, vp, i32, f32, f32, f32, f32, vp]),
    "macr_mf_trainer_create": (i32, [C.POINTER(vp), vp, vp, vp, i64, vp, vp, vp, i64] + [vp] * 6 +
                               [i32, i32, C.POINTER(HParams), vp]),
    "macr_mf_trainer_step": (i32, [vp, vp, vp, vp, i32, vp]),
    "macr_mf_trainer_step_host": (i32, [vp, vp, vp, vp, i32, vp]),
    "macr_mf_trainer_run": (i32, [vp, vp, i32, i32, vp]),
    "macr_mf_trainer_run_host": (i32, [vp, vp, i32, i32, vp]),
    "macr_mf_trainer_set_mode": (i32, [vp, i32]),
    "macr_mf_trainer_launches_per_step": (i32, [vp]),
    "macr_mf_trainer_steps_done": (i64, [vp]),
    "macr_mf_trainer_set_steps_done": (i32, [vp, i64]),
    "macr_mf_trainer_destroy": (i32, [vp]),
    "macr_spmm_csr": (i32, [vp, vp, vp, i64, vp, i32, vp, vp]),
    "macr_lgcn_propagate": (i32, [vp, vp, vp, vp, i64, vp, i64, i32, i32, vp, vp, vp]),
    "macr_spmm_plan_create": (i32, [vp, i64, C.POINTER(vp)]),
    "macr_spmm_plan_destroy": (i32, [vp]),
    "macr_spmm_csr_planned": (i32, [vp, vp, vp, vp, i64, vp, i32, vp, vp]),
    "macr_lgcn_propagate_planned": (i32, [vp, vp, vp, vp, vp, i64, vp, i64, i32, i32, vp, vp, vp]),
    "macr_lgcn_trainer_create": (i32, [C.POINTER(vp), vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i64] +
                                 [vp] * 6 + [i32, i32, i32, C.POINTER(HParams), vp]),
    "macr_lgcn_trainer_step": (i32, [vp, vp, vp, vp, i32, i32, vp]),
    "macr_lgcn_trainer_step_host": (i32, [vp, vp, vp, vp, i32, i32, vp]),
    "macr_lgcn_trainer_run": (i32, [vp, vp, i32, i32, i32, vp]),
    "macr_lgcn_trainer_run_host": (i32, [vp, vp, i32, i32, i32, vp]),
    "macr_lgcn_trainer_set_mode": (i32, [vp, i32]),
    "macr_lgcn_trainer_embeddings": (i32, [vp, C.POINTER(vp)]),
    "macr_lgcn_trainer_launches_per_step": (i32, [vp]),
    "macr_lgcn_trainer_steps_done": (i64, [vp]),
    "macr_lgcn_trainer_set_steps_done": (i32, [vp, i64]),
    "macr_lgcn_trainer_destroy": (i32, [vp]),
    "macr_score_gates": (i32, [vp, i64, i32, vp, vp, vp]),
    "macr_gather_rows": (i32, [vp, vp, i32, i32, vp, vp]),
    "macr_score_topk_workspace_bytes": (sz, [i32, i64, i32]),
    "macr_score_topk": (i32, [vp, i32, vp, i64, i32, vp, vp, f32, vp, vp, i32, C.c_int32, vp, vp,
                              vp, sz, vp]),
    "macr_score_topk_tc_workspace_bytes": (sz, [i32, i64, i32]),
    "macr_score_topk_tc": (i32, [vp, i32, vp, i64, i32, vp, vp, f32, vp, vp, i32, C.c_int32, vp, vp,
                                 vp, sz, vp, vp]),
    "macr_score_tc_items_bytes": (sz, [i64]),
    "macr_score_tc_prepare_items": (i32, [vp, i64, i32, vp, f32, vp, sz, vp]),
    "macr_score_topk_tc_prepared_workspace_bytes": (sz, [i32, i64, i32]),
    "macr_score_topk_tc_prepared": (i32, [vp, i32, vp, i64, i32, vp, vp, f32, vp, vp, i32, C.c_int32,
                                          vp, vp, vp, vp, sz, vp, vp]),
    "macr_score_matrix": (i32, [vp, i32, vp, i64, i32, vp, vp, f32, vp, vp]),
    "macr_topk_merge": (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    "macr_topk_rows": (i32, [vp, i32, i32, i32, vp, vp]),
    "macr_foldout_metrics": (i32, [vp, i32, i32, vp, vp, vp, vp, vp]),
    "macr_sample_mf": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp]),
    "macr_sample_lgcn": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp]),
    "macr_pairset_build": (i32, [vp, vp, i32, vp, i32]),
    "macr_sample_mf_epoch": (i32, [vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, i32, i32, vp]),
    "macr_sample_lgcn_epoch": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "macr_shard_pack": (i32, [vp, vp, C.POINTER(ShardDesc), vp, i32, i32, vp, vp, vp]),
    "macr_shard_unpack": (i32, [vp, vp, C.POINTER(ShardDesc), vp, i32, i32, vp]),
    "macr_shard_push": (i32, [vp, vp, C.POINTER(ShardDesc), vp, i32, i32, vp, C.POINTER(vp),
                              C.POINTER(vp), vp]),
    "macr_shard_barrier": (i32, [C.POINTER(vp), i32, i32, C.c_uint64, vp, vp]),
    "macr_mf_trainer_ipc_export": (i32, [vp, C.c_char_p]),
    "macr_mf_trainer_shard": (i32, [vp, C.POINTER(ShardDesc)] + [C.POINTER(vp)] * 3),
    "macr_mf_trainer_peer_error": (i32, [vp, C.POINTER(i32)]),
    "macr_lgcn_trainer_ipc_export": (i32, [vp, C.c_char_p]),
    "macr_lgcn_trainer_shard": (i32, [vp, C.POINTER(ShardDesc)] + [C.POINTER(vp)] * 5),
    "macr_lgcn_trainer_peer_error": (i32, [vp, C.POINTER(i32)]),
    "macr_topk_merge_peers": (i32, [C.POINTER(vp)] * 4 + [i32, i32, i32, i32, vp]),
    "macr_ipc_alloc": (i32, [sz, C.POINTER(vp), C.c_char_p]),
    "macr_ipc_open": (i32, [C.c_char_p, C.POINTER(vp)]),
    "macr_ipc_close": (i32, [vp]),
    "macr_ipc_free": (i32, [vp]),
}


def lib():
    """Load libmacr_b200.so (built by ``__graft_entry__.build()`` / ``make -C macr_b200/csrc``)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MacrError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc, what=""):
    if rc != 0:
        msg = lib().macr_last_error()
        raise MacrError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
