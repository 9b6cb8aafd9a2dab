"""ctypes binding of oracle/_ref/libmacr_ref_eval.so -- the REFERENCE's own C++ evaluator
(compiled from /root/reference by oracle/Makefile).  TEST INFRASTRUCTURE ONLY.
Same call contract as macr_lightgcn/evaluator/cpp/evaluate_foldout.py:12-18."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libmacr_ref_eval.so")
_LIB = None


def available():
    return os.path.exists(_SO)


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_SO)
    return _LIB


def top_k_array_index(scores, top_k, thread_num=None):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    rows, cols = scores.shape
    thread_num = thread_num or (os.cpu_count() or 1) * 5
    out = np.zeros((rows, top_k), np.int32)
    _lib().ref_c_top_k_array_index(scores.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(cols),
                                   C.c_int(rows), C.c_int(top_k), C.c_int(thread_num),
                                   out.ctypes.data_as(C.POINTER(C.c_int)))
    return out


def evaluate_foldout(rankings, test_items, thread_num=None):
    rankings = np.ascontiguousarray(rankings, dtype=np.int32)
    users, K = rankings.shape
    thread_num = thread_num or (os.cpu_count() or 1) * 5
    truths = [np.ascontiguousarray(t, dtype=np.int32) for t in test_items]
    ptrs = (C.POINTER(C.c_int) * users)(*[t.ctypes.data_as(C.POINTER(C.c_int)) for t in truths])
    nums = np.array([len(t) for t in truths], np.int32)
    out = np.zeros((users, 5 * K), np.float32)
    _lib().ref_evaluate_foldout(C.c_int(users), rankings.ctypes.data_as(C.POINTER(C.c_int)),
                                C.c_int(K), ptrs, nums.ctypes.data_as(C.POINTER(C.c_int)),
                                C.c_int(thread_num), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def eval_score_matrix_foldout(score_matrix, test_items, top_k=20, thread_num=None):
    if len(score_matrix) != len(test_items):
        raise ValueError("The lengths of score_matrix and test_items are not equal.")
    rk = top_k_array_index(score_matrix, top_k, thread_num)
    return evaluate_foldout(rk, test_items, thread_num)
