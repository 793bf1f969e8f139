"""Oracle (test infrastructure): ctypes binding of oracle/csrc/oracle.c (true fmaf() scoring + exact masked top-N).
Built by `make -C oracle` (also run by __graft_entry__.build()); tests skip the C comparison when gcc is missing."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


def load():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "csrc", "oracle.c")):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(LIB)
    f32, i32, i64 = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    lib.oracle_scores_fma32.argtypes = [f32, f32, C.c_int64, C.c_int64, C.c_int, f32]
    lib.oracle_topn_exact.argtypes = [f32, f32, C.c_int64, C.c_int, i32, C.c_int64, C.c_int, i64, i32, i32, f32]
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def scores_fma32(P_rows, Q):
    P_rows, Q = np.ascontiguousarray(P_rows, np.float32), np.ascontiguousarray(Q, np.float32)
    out = np.empty((P_rows.shape[0], Q.shape[0]), np.float32)
    load().oracle_scores_fma32(_p(P_rows, C.c_float), _p(Q, C.c_float), P_rows.shape[0], Q.shape[0], P_rows.shape[1], _p(out, C.c_float))
    return out


def topn_exact(P, Q, users, N, uq_indptr, uq_items):
    P, Q = np.ascontiguousarray(P, np.float32), np.ascontiguousarray(Q, np.float32)
    users = np.ascontiguousarray(users, np.int32)
    uq_indptr, uq_items = np.ascontiguousarray(uq_indptr, np.int64), np.ascontiguousarray(uq_items, np.int32)
    ids, sc = np.empty((len(users), N), np.int32), np.empty((len(users), N), np.float32)
    load().oracle_topn_exact(_p(P, C.c_float), _p(Q, C.c_float), Q.shape[0], P.shape[1], _p(users, C.c_int32), len(users), N,
                             _p(uq_indptr, C.c_int64), _p(uq_items, C.c_int32), _p(ids, C.c_int32), _p(sc, C.c_float))
    return ids, sc
