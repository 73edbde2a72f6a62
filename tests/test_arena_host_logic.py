"""The arena's HOST logic (game bookkeeping, result assignment, PGN and sample export, game groups) with the device
search replaced by a stand-in that plays random legal moves -- so that this logic is covered without a GPU
(the real search under the same arena is tests/test_selfplay_gpu.py)."""
import numpy as np

import crazyara_b200.selfplay as sp
from crazyara_b200.export import BLACK_WIN, DRAWN, WHITE_WIN, read_dataset


class FakeAgent:
    """MCTSAgent's surface as the arena uses it; 'search' = uniform policy over the legal moves, seeded best move."""
    created = []

    best_move_q = 0.25  # EvalInfo::bestMoveQ[0] of every stand-in search

    def __init__(self, net, settings, device=0, n_trees=1, max_nodes=0):
        self.n_trees, self.states, self.applied = n_trees, [None] * n_trees, []
        self.rng = np.random.default_rng(len(FakeAgent.created))
        self.seed, self.limits = int(settings.seed), []
        FakeAgent.created.append(self)

    def set_search_limits(self, simulations, nodes, tree=-1):
        self.limits.append((tree, simulations, nodes))

    def set_position(self, state, tree=0):
        self.states[tree] = state

    def evaluate_board_state(self, state=None):
        self.res = []
        for st in self.states:
            moves = [st.action_to_uci(a) for a in st.legal_actions()]
            k = len(moves)
            pol = np.full(k, 1.0 / k) if k else np.zeros(0)
            self.res.append(dict(moves=moves, policy=pol, q=np.full(k, -0.5, np.float32), best_idx=int(self.rng.integers(k)) if k else -1,
                                 nodes=10, nodes_pre_search=0, best_move_q=FakeAgent.best_move_q))

    def result(self, tree=0):
        return self.res[tree]

    def last_go_ms(self):
        return 1.0

    def apply_move_to_tree(self, move, tree=0):
        self.applied.append((tree, move))

    def close(self):
        pass


def _arena(monkeypatch, tmp_path, **kw):
    FakeAgent.created = []
    monkeypatch.setattr(sp, "MCTSAgent", FakeAgent)
    monkeypatch.setattr(sp, "encode_planes", lambda boards, mode, version, normalize=False: np.ones((len(boards), 34, 8, 8), np.float32))
    st = sp.rl_settings("crazyhouse")
    return sp.Arena(kw.pop("net", None), st, variant=1, n_games=kw.pop("n_games", 3), temperature_moves=0, max_plies=kw.pop("max_plies", 60),
                    seed=4, **kw)


def test_games_results_pgn_and_samples(monkeypatch, tmp_path):
    from crazyara_b200.export import TrainDataExporter
    ex = TrainDataExporter(str(tmp_path / "d.zarr"), "crazyhouse", channels=34, number_chunks=40, chunk_size=16)
    pgn = str(tmp_path / "games.pgn")
    arena = _arena(monkeypatch, tmp_path, exporter=ex, pgn_path=pgn, reuse_tree=True)
    res = arena.run(min_games=6, max_steps=400)
    assert res["games"] >= 6 and res["nodes"] == 10 * res["moves"]
    # every played move was announced to the tree first (reuse_tree), one call per game and step
    assert len(FakeAgent.created[0].applied) == res["moves"]
    # one PGN record per finished game, result token consistent with the adjudication
    text = open(pgn).read()
    blocks = text.strip().split("\n\n\n")
    assert len(blocks) == res["games"]
    results = []
    for block, (plies, term, stm) in zip(blocks, arena.finished):
        header, body = block.split("\n\n", 1)
        assert f'[PlyCount "{plies}"]' in header and '[Variant "crazyhouse"]' in header
        token = body.split()[-1]
        expect = "1/2-1/2" if term not in (0, 2) else (("0-1" if stm == 0 else "1-0") if term == 0 else ("1-0" if stm == 0 else "0-1"))
        assert token == expect and f'[Result "{token}"]' in header
        if term == 0:
            assert body.split()[-2].endswith("#")          # the game ended by mate (or a variant loss): '#' on the last move
        results.append({"1-0": WHITE_WIN, "0-1": BLACK_WIN}.get(token, DRAWN))
    # the exported samples: one per searched position of the finished games, value = result from the mover's view
    starts = read_dataset(str(tmp_path / "d.zarr"), "start_indices")
    n_games = ex.game_idx
    assert n_games == res["games"] and list(starts[1:n_games + 1] - starts[:n_games]) == [p for p, _, _ in arena.finished]
    values = read_dataset(str(tmp_path / "d.zarr"), "y_value")
    plys = read_dataset(str(tmp_path / "d.zarr"), "plys_to_end")
    for g, result in enumerate(results):
        v = values[starts[g]:starts[g + 1]]
        white_to_move = (np.arange(len(v)) % 2 == 0)     # games start with white to move
        want = 0 if result == DRAWN else (1 if result == WHITE_WIN else -1)
        assert np.array_equal(v, np.where(white_to_move, want, -want))
        assert list(plys[starts[g]:starts[g + 1]]) == list(range(len(v), 0, -1))
    pol = read_dataset(str(tmp_path / "d.zarr"), "y_policy")[:starts[n_games]]
    assert np.allclose(pol.sum(axis=1), 1.0, atol=1e-5)


def test_reference_self_play_details(monkeypatch, tmp_path):
    """What SelfPlay::generate_game does around the search (rl/selfplay.cpp:192-261): the exported q is bestMoveQ[0] (not
    the sampled move's edge Q), the node budget is jittered per search, every game group draws its own Dirichlet seed,
    the temperature decays with the ply, low policy entries are clipped before export, games can be resigned."""
    from crazyara_b200.export import TrainDataExporter
    ex = TrainDataExporter(str(tmp_path / "d.zarr"), "crazyhouse", channels=34, number_chunks=40, chunk_size=16)
    arena = _arena(monkeypatch, tmp_path, exporter=ex, net=[None, None], n_games=4, max_plies=20)
    res = arena.run(min_games=4, max_steps=200)
    n = read_dataset(str(tmp_path / "d.zarr"), "start_indices")[ex.game_idx]
    q = read_dataset(str(tmp_path / "d.zarr"), "y_best_move_q")[:n]
    assert n > 0 and np.allclose(q, FakeAgent.best_move_q)            # not the -0.5 of the per-move Q vector
    assert FakeAgent.created[0].seed != FakeAgent.created[1].seed     # groups (and processes, via `seed`) differ
    lim = [nodes for a in FakeAgent.created for (_, _, nodes) in a.limits]
    assert len(lim) >= res["moves"] and min(lim) >= 760 and max(lim) <= 840 and len(set(lim)) > 3   # 800 +- 5 %
    # temperature decay (get_current_temperature): ply 0 flattens with T = 0.8, ply 10 sharpens with T = 0.8 * 0.92^10
    arena.temperature_moves, arena.temperature = 15, 0.8
    pol = np.array([0.7, 0.2, 0.1])
    picks0 = [arena._pick(dict(policy=pol, best_idx=0), 0) for _ in range(4000)]
    picks10 = [arena._pick(dict(policy=pol, best_idx=0), 10) for _ in range(4000)]
    p0 = pol ** (1 / 0.8) / (pol ** (1 / 0.8)).sum()
    t10 = 0.8 * 0.92 ** 10
    p10 = pol ** (1 / t10) / (pol ** (1 / t10)).sum()
    assert abs(np.mean(np.array(picks0) == 0) - p0[0]) < 0.03 and abs(np.mean(np.array(picks10) == 0) - p10[0]) < 0.03
    assert arena._pick(dict(policy=pol, best_idx=2), 15) == 2         # after Temperature_Moves: the best move
    # sharpen_distribution (blazeutil.h:94-105) and apply_quantile_clipping (agent.cpp:118-127)
    assert np.allclose(sp.sharpen_distribution(np.array([0.5, 0.495, 0.005]), 0.01), [0.5 / 0.995, 0.495 / 0.995, 0.0])
    assert np.allclose(sp.sharpen_distribution(np.array([0.004, 0.003]), 0.01), [0.004, 0.003])   # max below thresh: untouched
    assert np.allclose(sp.quantile_clip(np.array([0.6, 0.3, 0.06, 0.04]), 0.25), [0.6 / 0.9, 0.3 / 0.9, 0, 0])
    # resignation: the mover's bestMoveQ below the threshold ends the game in favour of the side then to move
    FakeAgent.best_move_q = -0.95
    try:
        arena2 = _arena(monkeypatch, tmp_path, n_games=2, max_plies=50, resign_probability=1.0)
        arena2.run(max_steps=1)
        assert arena2.resigned == 2 and [f[:2] for f in arena2.finished] == [(1, sp.TERMINAL_WIN)] * 2
        assert all(f[2] == 1 for f in arena2.finished)                # black to move after white's first move: black wins
    finally:
        FakeAgent.best_move_q = 0.25


def test_mixed_variant_games(monkeypatch, tmp_path):
    """BASELINE cfg 5: King of the Hill and Three-check side by side (one variant per game, cycled); a finished game is
    replaced by a new game of the same variant."""
    FakeAgent.created = []
    monkeypatch.setattr(sp, "MCTSAgent", FakeAgent)
    st = sp.rl_settings("lichess")
    arena = sp.Arena(None, st, variant=[2, 3], n_games=4, temperature_moves=0, max_plies=6, seed=1, resign_probability=0.0)
    assert [s.variant for s in arena.states] == [2, 3, 2, 3]
    arena.run(max_steps=7)   # every game is adjudicated at 6 plies and restarted
    assert len(arena.finished) >= 4 and [s.variant for s in arena.states] == [2, 3, 2, 3]


def test_game_groups_split_the_games(monkeypatch, tmp_path):
    arena = _arena(monkeypatch, tmp_path, net=[None, None], n_games=4, max_plies=12)
    assert [a.n_trees for a in FakeAgent.created] == [2, 2]
    res = arena.run(max_steps=30)
    assert res["moves"] == 4 * 30 and res["games"] >= 4 and all(p <= 12 for p, _, _ in arena.finished)
    assert all(s is not None for a in FakeAgent.created for s in a.states)


def test_launcher_worker_end_to_end(monkeypatch, tmp_path, capsys):
    """python -m crazyara_b200.selfplay: one worker's whole run (network buffers, arena, export, PGN) with the device
    pieces replaced by stand-ins."""
    import argparse
    import crazyara_b200.nn as nn_mod

    class FakeNet:
        def __init__(self, ctx, device, batch, path):
            self.batch = batch

        def get_nb_input_values_total(self):
            return 34 * 64

        def close(self):
            pass

    FakeAgent.created = []
    monkeypatch.setattr(sp, "MCTSAgent", FakeAgent)
    monkeypatch.setattr(sp, "encode_planes", lambda boards, mode, version, normalize=False: np.ones((len(boards), 34, 8, 8), np.float32))
    monkeypatch.setattr(nn_mod, "NeuralNetAPI", FakeNet)
    plan = sp.plan_workers(8, [0, 1], str(tmp_path))
    args = argparse.Namespace(model="m.arab", mode="crazyhouse", variant="1", chess960=False, input_version=1, batch_size=8,
                              nodes=800, max_plies=16, chunks=4, seed=1, export=True, pgn=True, games_per_worker=4, seconds=30.0)
    sp._worker(plan[1], args)
    out = capsys.readouterr().out
    assert "data_gpu_1.zarr" in out and "games/h" in out
    assert [a.n_trees for a in FakeAgent.created] == [2, 2]          # 4 games of this worker in two groups
    assert open(plan[1]["pgn"]).read().count("[Event ") >= 4
    assert read_dataset(plan[1]["zarr"], "start_indices")[1] > 0
