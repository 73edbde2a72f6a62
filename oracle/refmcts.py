"""ctypes wrapper of oracle/_ref/libref_mcts.so: the REFERENCE's own search code (node.cpp, searchthread.cpp, MCTSAgent,
... compiled unchanged by `make -C oracle ref`) driven through MCTSAgent::evaluate_board_state.  Test infrastructure."""
import ctypes
import os

import numpy as np

from .search import OSettings

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_mcts.so")


class RefResult(ctypes.Structure):
    _fields_ = [("n_moves", ctypes.c_int), ("no_visit_idx", ctypes.c_int), ("best_idx", ctypes.c_int), ("pv_len", ctypes.c_int),
                ("root_value", ctypes.c_float), ("best_move_q", ctypes.c_float), ("visit_sum", ctypes.c_uint),
                ("free_visits", ctypes.c_uint), ("nodes", ctypes.c_uint), ("moves", ctypes.c_uint * 512),
                ("visits", ctypes.c_uint * 512), ("q", ctypes.c_float * 512), ("prior", ctypes.c_float * 512),
                ("policy", ctypes.c_double * 512), ("pv", ctypes.c_uint * 256)]


NET_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_ulonglong),
                          ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float))
_LIB = None


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(PATH)
        _LIB.ref_mcts_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p),
                                      ctypes.c_int, ctypes.POINTER(OSettings), ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.POINTER(RefResult)]
    return _LIB


def run(pos_for_uci, fen, variant_id, is960, premoves, settings, net_fn=None, channels=None, n_labels=None):
    """One MCTSAgent::evaluate_board_state.  net_fn(planes[n,C,8,8], keys[n]) -> (value[n], prob[n,L]) (keys = the
    positions' Zobrist keys, for hash-derived stand-in networks), or None for oracle/fake.c.  pos_for_uci: an oracle
    Position used only to spell the root moves as UCI strings."""
    L = lib()
    cb = None
    if net_fn is not None:
        def _cb(_ctx, planes, _keys, n, value, prob):
            if n <= 0:
                return
            x = np.ctypeslib.as_array(planes, shape=(n, channels, 8, 8)).copy()
            v, p = net_fn(x, np.ctypeslib.as_array(_keys, shape=(n,)).copy())
            np.ctypeslib.as_array(value, shape=(n,))[:] = np.asarray(v, np.float32)
            np.ctypeslib.as_array(prob, shape=(n, n_labels))[:] = np.asarray(p, np.float32)
        cb = NET_FN(_cb)
    arr = (ctypes.c_char_p * max(len(premoves), 1))(*[m.encode() for m in premoves])
    out = RefResult()
    rc = L.ref_mcts_run((fen or "").encode(), variant_id, int(is960), arr, len(premoves), ctypes.byref(settings),
                        ctypes.cast(cb, ctypes.c_void_p) if cb else None, None, ctypes.byref(out))
    assert rc == 0
    k = out.n_moves
    return dict(moves=[pos_for_uci.uci(int(m)) for m in out.moves[:k]], visits=np.array(out.visits[:k], np.uint32),
                q=np.array(out.q[:k], np.float32), prior=np.array(out.prior[:k], np.float32),
                policy=np.array(out.policy[:k], np.float64), root_value=float(np.float32(out.root_value)),
                best_move_q=float(np.float32(out.best_move_q)), visit_sum=out.visit_sum, free_visits=out.free_visits,
                nodes=out.nodes, best_idx=out.best_idx, pv_len=out.pv_len, no_visit_idx=out.no_visit_idx)
