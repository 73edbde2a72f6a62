"""Python driver of the C search oracle (oracle/mcts.c) -- TEST INFRASTRUCTURE ONLY.

run_search() mirrors MCTSAgent::evaluate_board_state (agents/mctsagent.cpp:292-337) with Threads=1: the network is
any callable planes[n,C,8,8] -> (value[n], prob[n,L]) so the same search can be driven by the fp32 torch oracle
(CPU baseline), by the GPU network through its host API (search parity) or by the hash-derived fake backend."""
import ctypes

import numpy as np

from .chess import Position, lib

MODES = {"crazyhouse": 0, "chess": 1, "lichess": 2}


class OSettings(ctypes.Structure):
    _fields_ = [("batch_size", ctypes.c_int), ("dirichlet_epsilon", ctypes.c_float), ("dirichlet_alpha", ctypes.c_float),
                ("node_policy_temperature", ctypes.c_float), ("q_value_weight", ctypes.c_float),
                ("q_veto_delta", ctypes.c_float), ("cpuct_init", ctypes.c_float), ("cpuct_base", ctypes.c_float),
                ("mcts_solver", ctypes.c_int), ("virtual_style", ctypes.c_int), ("virtual_mix_threshold", ctypes.c_uint),
                ("simulations", ctypes.c_uint), ("nodes", ctypes.c_uint), ("seed", ctypes.c_ulonglong),
                ("mode", ctypes.c_int), ("input_version", ctypes.c_int), ("threads", ctypes.c_int),
                ("epsilon_greedy_counter", ctypes.c_int), ("epsilon_checks_counter", ctypes.c_int), ("reserved", ctypes.c_int)]


def default_settings(mode, **kw):
    L = _lib()
    s = OSettings()
    L.osettings_default(ctypes.byref(s), MODES[mode] if isinstance(mode, str) else mode)
    for k, v in kw.items():
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


_SET = False


def _lib():
    global _SET
    L = lib()
    if not _SET:
        L.osearch_new.restype = ctypes.c_void_p
        L.osearch_new.argtypes = [ctypes.c_void_p]
        for name in ("osearch_free", "osearch_channels", "osearch_nb_labels", "osearch_create_mini_batch",
                     "osearch_continue", "osearch_root_num_children", "osearch_root_no_visit_idx",
                     "osearch_root_visits", "osearch_root_free_visits", "osearch_best_move_idx",
                     "osearch_root_node_type"):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.osearch_set_root.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.osearch_apply_move.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        L.osearch_root_reused.argtypes = [ctypes.c_void_p]
        L.osearch_root_results.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.osearch_apply_results.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.osearch_planes.restype = ctypes.c_void_p
        L.osearch_planes.argtypes = [ctypes.c_void_p]
        L.osearch_root_stats.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 5
        L.osearch_root_value.restype = ctypes.c_float
        L.osearch_root_value.argtypes = [ctypes.c_void_p]
        L.osearch_best_move_q.restype = ctypes.c_float
        L.osearch_best_move_q.argtypes = [ctypes.c_void_p]
        L.osearch_pv.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        for name in ("osearch_num_nodes", "osearch_sum_select_k", "osearch_sum_depth"):
            getattr(L, name).restype = ctypes.c_ulonglong
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.osearch_create_mini_batch_t.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.osearch_apply_results_t.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.osearch_planes_t.restype = ctypes.c_void_p
        L.osearch_planes_t.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.osearch_batch_keys_t.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.oglibc_rand_sequence.argtypes = [ctypes.c_uint, ctypes.c_int, ctypes.c_void_p]
        L.odirichlet_noise.argtypes = [ctypes.c_ulonglong, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
        L.osearch_batch_keys.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ofake_eval.argtypes = [ctypes.c_ulonglong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _SET = True
    return L


class Search:
    def __init__(self, settings):
        self.L = _lib()
        self.settings = settings
        self.h = self.L.osearch_new(ctypes.byref(settings))
        self.channels = self.L.osearch_channels(self.h)
        self.n_labels = self.L.osearch_nb_labels(self.h)
        self.batch = settings.batch_size

    def close(self):
        if self.h:
            self.L.osearch_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _planes(self, n):
        ptr = self.L.osearch_planes(self.h)
        arr = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(self.batch, self.channels, 8, 8))
        return arr[:n]

    def _keys(self, n):
        k = np.zeros(max(n, 1), np.uint64)
        self.L.osearch_batch_keys(self.h, k.ctypes.data)
        return k[:n]

    def _eval_thread(self, t, n, net_fn, with_keys):
        """network call for the batch of logical search thread t: planes (and keys) of its n new leaves"""
        ptr = self.L.osearch_planes_t(self.h, t)
        planes = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)),
                                       shape=(self.batch, self.channels, 8, 8))[:n].copy()
        if not with_keys:
            v, p = net_fn(planes)
        else:
            k = np.zeros(max(n, 1), np.uint64)
            self.L.osearch_batch_keys_t(self.h, t, k.ctypes.data)
            v, p = net_fn(planes, k[:n])
        return np.ascontiguousarray(v, np.float32), np.ascontiguousarray(p, np.float32)

    def _run_two_threads(self, net_fn, max_iterations, with_keys):
        """Threads = 2 in the fixed schedule of oracle/mcts.h: each logical thread is run_search_thread
        (searchthread.cpp:418-426) -- while the loop condition holds: create_mini_batch, predict, back up -- and the two
        alternate phase by phase, thread 1 half an iteration behind thread 0."""
        L, h = self.L, self.h
        pending = [None, None]   # per thread: number of new leaves of the batch waiting for its backup, None = no batch
        alive = [True, True]
        iters = evals = 0

        def sel(t):
            nonlocal iters, evals
            if not (alive[t] and L.osearch_continue(h) and iters < max_iterations):
                alive[t] = False
                return
            pending[t] = L.osearch_create_mini_batch_t(h, t)
            iters += 1
            evals += pending[t]

        def bk(t):
            if pending[t] is None:
                return
            n = pending[t]
            if n > 0:
                v, p = self._eval_thread(t, n, net_fn, with_keys)
            else:
                v, p = np.zeros(1, np.float32), np.zeros(1, np.float32)
            L.osearch_apply_results_t(h, t, v.ctypes.data, p.ctypes.data)
            pending[t] = None
        sel(0)
        sel(1)
        while pending[0] is not None or pending[1] is not None:
            for t in (0, 1):
                bk(t)
                sel(t)
        return iters, evals

    def run(self, pos, net_fn, max_iterations=1 << 30, with_keys=False, threads=1):
        """Returns the result dict of update_eval_info (evalinfo.cpp:195-249) plus counters."""
        L, h = self.L, self.h
        self.pos = pos
        iters = 0
        evals = 0
        rc = L.osearch_set_root(h, pos._buf)
        self.reused = rc == 2
        self.nodes_pre_search = 0
        if rc:
            if rc == 1:
                v, p = net_fn(self._planes(1).copy(), self._keys(1)) if with_keys else net_fn(self._planes(1).copy())
                v = np.ascontiguousarray(v, np.float32)
                p = np.ascontiguousarray(p, np.float32)
                L.osearch_root_results(h, v.ctypes.data, p.ctypes.data)
                evals += 1
            else:
                self.nodes_pre_search = L.osearch_root_visits(h) - L.osearch_root_free_visits(h)
                L.osearch_root_reused(h)
            if L.osearch_root_num_children(h) > 1 and threads == 2:
                it2, ev2 = self._run_two_threads(net_fn, max_iterations, with_keys)
                iters += it2
                evals += ev2
            elif L.osearch_root_num_children(h) > 1:
                while L.osearch_continue(h) and iters < max_iterations:
                    n = L.osearch_create_mini_batch(h)
                    if n > 0:
                        v, p = (net_fn(self._planes(n).copy(), self._keys(n)) if with_keys
                                else net_fn(self._planes(n).copy()))
                        v = np.ascontiguousarray(v, np.float32)
                        p = np.ascontiguousarray(p, np.float32)
                        evals += n
                    else:
                        v = np.zeros(1, np.float32)
                        p = np.zeros(1, np.float32)
                    L.osearch_apply_results(h, v.ctypes.data, p.ctypes.data)
                    iters += 1
        return self.result(iters, evals)

    def apply_move(self, move):
        """MCTSAgent::apply_move_to_tree: keep the subtree behind `move` (oracle move code) for the next run()."""
        return bool(self.L.osearch_apply_move(self.h, int(move)))

    def result(self, iters=0, evals=0):
        L, h = self.L, self.h
        k = L.osearch_root_num_children(h)
        moves = np.zeros(max(k, 1), np.uint32)
        visits = np.zeros(max(k, 1), np.uint32)
        q = np.zeros(max(k, 1), np.float32)
        prior = np.zeros(max(k, 1), np.float32)
        pol = np.zeros(max(k, 1), np.float64)
        if k > 0:
            L.osearch_root_stats(h, moves.ctypes.data, visits.ctypes.data, q.ctypes.data, prior.ctypes.data, pol.ctypes.data)
        pv = np.zeros(256, np.uint32)
        npv = L.osearch_pv(h, pv.ctypes.data, 256) if k > 0 and L.osearch_root_no_visit_idx(h) > 0 else 0
        res = dict(moves=[self.pos.uci(int(m)) for m in moves[:k]], visits=visits[:k].copy(), q=q[:k].copy(),
                   prior=prior[:k].copy(), policy=pol[:k].copy(), root_value=L.osearch_root_value(h) if k > 0 else 0.0,
                   visit_sum=L.osearch_root_visits(h), free_visits=L.osearch_root_free_visits(h),
                   nodes=L.osearch_root_visits(h) - L.osearch_root_free_visits(h),
                   best_idx=L.osearch_best_move_idx(h) if k > 0 else -1,
                   best_move_q=L.osearch_best_move_q(h) if k > 0 else 0.0, node_type=L.osearch_root_node_type(h),
                   pv_len=npv, iterations=iters, evals=evals, tree_nodes=L.osearch_num_nodes(h),
                   sum_select_k=L.osearch_sum_select_k(h), sum_depth=L.osearch_sum_depth(h))
        if k > 0:
            res["best_move"] = res["moves"][res["best_idx"]]
        return res


def dirichlet_noise(seed, n, alpha):
    out = np.zeros(n, np.float32)
    _lib().odirichlet_noise(seed, n, alpha, out.ctypes.data)
    return out


def fake_net(n_labels):
    """Hash-derived backend (oracle/fake.c): net_fn(planes, keys) for Search.run(..., with_keys=True)."""
    L = _lib()

    def fn(planes, keys):
        n = len(keys)
        v = np.zeros(n, np.float32)
        p = np.zeros((n, n_labels), np.float32)
        for i, k in enumerate(keys):
            L.ofake_eval(int(k), n_labels, v[i:].ctypes.data, p[i].ctypes.data)
        return v, p
    return fn


def hash_net(n_labels):
    """A second hash-derived stand-in network, WITHOUT ties: value as oracle/fake.c, priors (1 + m/2^23) * 2^(-9-t) with
    23 hashed mantissa bits m and t in 0..7.  oracle/fake.c's priors take only 2048 distinct values, so legal moves tie
    all the time and the order std::sort leaves tied priors in (node.cpp:464-470: unstable, unspecified) would decide
    visit orders; comparisons against the reference's compiled search (oracle/refmcts.py) use this one.
    net_fn(planes, keys) for Search.run(..., with_keys=True)."""
    m1, m2 = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)

    def zmix(z):
        z = (z ^ (z >> np.uint64(30))) * m1
        z = (z ^ (z >> np.uint64(27))) * m2
        return z ^ (z >> np.uint64(31))
    idx = (np.arange(n_labels, dtype=np.uint64) + np.uint64(1)) * np.uint64(0xD6E8FEB86659FD93)

    def fn(planes, keys):
        keys = np.asarray(keys, np.uint64)
        with np.errstate(over="ignore"):
            h0 = zmix(keys ^ np.uint64(0x9E3779B97F4A7C15))
            h = zmix(keys[:, None] + idx[None, :])
        v = (((h0 >> np.uint64(11)) & np.uint64(0xFFFF)).astype(np.int64) - 32768).astype(np.float32) / np.float32(65536)
        t = (h & np.uint64(7)).astype(np.uint32)
        m = ((h >> np.uint64(20)) & np.uint64(0x7FFFFF)).astype(np.uint32)
        p = (((np.uint32(127 - 9) - t) << np.uint32(23)) | m).view(np.float32)
        return v, np.ascontiguousarray(p)
    return fn
