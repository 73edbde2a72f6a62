"""The fp32 select fast path (search_dev.cuh pick_fast) may only claim certainty when its winner is the exact
(reference-arithmetic) first maximum; everything else must fall back.  Checked on random, tied and near-tied children
against an independent numpy restatement of node.cpp:1056-1063 / :1150-1167."""
import ctypes

import numpy as np

from tests.hostemu import lib


def _pick(p, q, n, cput, vs):
    L = lib()
    out = (ctypes.c_int * 3)()
    p = np.ascontiguousarray(p, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    n = np.ascontiguousarray(n, np.uint32)
    L.he_pick_both.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                               ctypes.c_uint, ctypes.c_void_p]
    L.he_pick_both(len(p), p.ctypes.data, q.ctypes.data, n.ctypes.data, np.float32(cput), int(vs), out)
    return bool(out[0]), out[1], out[2]


def _reference_argmax(p, q, n, cput, vs):
    # blaze restructures (cput * P) * w into (P * w) * cput: all double, one final narrowing (oracle/mcts.c select_child_node)
    u = ((p.astype(np.float64) * (np.sqrt(np.float64(vs)) / (n.astype(np.float64) + 1.0))) * np.float64(np.float32(cput))).astype(np.float32)
    v = (q.astype(np.float32) + u).astype(np.float32)
    return int(np.argmax(v))  # first maximum


def test_fast_path_never_contradicts_exact_arithmetic():
    rng = np.random.default_rng(7)
    sure_count = total = 0
    for trial in range(4000):
        k = int(rng.integers(2, 60))
        p = rng.dirichlet(np.full(k, 0.3)).astype(np.float32)
        n = rng.integers(0, 2000, k).astype(np.uint32)
        q = rng.uniform(-1, 1, k).astype(np.float32)
        kind = trial % 5
        if kind == 1:      # exact duplicates
            j = int(rng.integers(1, k))
            p[j], q[j], n[j] = p[0], q[0], n[0]
        elif kind == 2:    # one-ulp neighbours
            j = int(rng.integers(1, k))
            p[j], n[j] = p[0], n[0]
            q[j] = np.nextafter(q[0], np.float32(2.0 if trial & 8 else -2.0))
        elif kind == 3:    # fresh children: Q_INIT everywhere, no visits
            q[:] = -1.0
            n[:] = 0
        elif kind == 4:    # disabled actions and huge visit counts
            q[int(rng.integers(0, k))] = np.float32(-2147483647.0)
            n[int(rng.integers(0, k))] = np.uint32(3_000_000_000)
        vs = int(n.astype(np.uint64).sum() % (1 << 32))
        cput = np.float32(2.5 + np.log((vs + 19652.0 + 1) / 19652.0))
        sure, fast_ci, exact_ci = _pick(p, q, n, cput, vs)
        assert exact_ci == _reference_argmax(p, q, n, cput, vs)
        if sure:
            assert fast_ci == exact_ci
            sure_count += 1
        total += 1
    assert sure_count > 0.5 * total  # the fast path must actually be taken most of the time


def test_exact_ties_are_never_certain():
    p = np.array([0.25, 0.25, 0.25, 0.25], np.float32)
    q = np.array([-1, -1, -1, -1], np.float32)
    n = np.zeros(4, np.uint32)
    sure, _, exact_ci = _pick(p, q, n, 2.5, 10)
    assert not sure and exact_ci == 0
