"""Self-play arena: concurrent games, one device tree each (games/hr leg of the metric)."""
import pytest


@pytest.mark.gpu
def test_arena_plays_legal_games_fake_backend():
    from crazyara_b200.selfplay import Arena, rl_settings
    st = rl_settings("crazyhouse", batch_size=8, nodes=60, simulations=240)
    arena = Arena(None, st, variant=1, n_games=6, temperature_moves=8, max_plies=40, seed=3)
    res = arena.run(min_games=6, max_steps=60)
    assert res["games"] >= 6 and res["moves"] > 0 and res["nodes"] > 0
    assert all(p <= 40 for p, _, _ in arena.finished)
    arena.close()


@pytest.mark.gpu
def test_arena_real_net_chess960(tmp_path):
    from crazyara_b200.nn import NeuralNetAPI
    from crazyara_b200.selfplay import Arena, rl_settings
    from crazyara_b200.weights import export_blob
    from oracle import net as onet
    arch = onet.arch_risev33(52, 76, True)
    blob = export_blob(onet.make_state_dict(arch, 0), arch, str(tmp_path / "v33.arab"), input_version=30)
    n_games, B = 8, 8
    net = NeuralNetAPI("gpu", 0, n_games * B, blob)
    st = rl_settings("chess", batch_size=B, nodes=100, simulations=400, input_version=3)
    arena = Arena(net, st, variant=0, n_games=n_games, is960=True, max_plies=12, seed=1)
    starts = {s.fen().split(" ")[0].split("/")[7] for s in arena.states}
    assert len(starts) > 1 and all(sorted(r) == sorted("RNBQKBNR") for r in starts)   # random chess960 set-ups
    res = arena.run(min_games=8, max_steps=14)
    assert res["games"] >= 8 and res["nps"] > 0
    arena.close()
    net.close()


@pytest.mark.gpu
def test_arena_with_tree_reuse():
    """Reuse_Tree in self-play: the subtree of the played move carries its visits into the next search."""
    from crazyara_b200.selfplay import Arena, rl_settings
    st = rl_settings("crazyhouse", batch_size=8, nodes=80, simulations=320)
    arena = Arena(None, st, variant=1, n_games=4, temperature_moves=6, max_plies=30, seed=5, reuse_tree=True)
    res = arena.run(min_games=4, max_steps=40)
    assert res["games"] >= 4 and res["reused_nodes"] > 0 and res["nodes"] > 0
    arena.close()


@pytest.mark.gpu
def test_arena_writes_pgn(tmp_path):
    """games.pgn of the reference's self-play: every finished game is appended with its result and SAN moves."""
    from crazyara_b200.selfplay import Arena, rl_settings
    st = rl_settings("crazyhouse", batch_size=8, nodes=60, simulations=240)
    path = str(tmp_path / "games.pgn")
    arena = Arena(None, st, variant=1, n_games=4, temperature_moves=8, max_plies=24, seed=11, pgn_path=path)
    res = arena.run(min_games=4, max_steps=40)
    arena.close()
    text = open(path).read()
    games = text.count('[Variant "crazyhouse"]')
    assert games == res["games"] >= 4
    assert text.count("[Result ") == games and '[PlyCount "0"]' not in text
    for block in text.strip().split("\n\n\n"):
        header, body = block.split("\n\n", 1)
        plies = int(header.split('[PlyCount "')[1].split('"')[0])
        tokens = [t for t in body.split() if not t.endswith(".")]
        assert len(tokens) == plies + 1 and tokens[-1] in ("1-0", "0-1", "1/2-1/2")


@pytest.mark.gpu
def test_arena_game_groups_on_separate_threads():
    """Two groups of games with an agent each, searched concurrently from two host threads.  Trees never interact, so the
    games are reproducible whatever the interleaving of the two host threads: two such arenas with the same seed play the
    same games.  (Every tree and every group draws its own Dirichlet noise -- seeds derived from the arena's -- so a
    one-group arena plays different games.)"""
    from crazyara_b200.selfplay import Arena, rl_settings
    st = rl_settings("crazyhouse", batch_size=8, nodes=60, simulations=240)
    kw = dict(variant=1, n_games=4, temperature_moves=0, max_plies=20, seed=2, node_random_factor=0.0)
    a2 = Arena([None, None], st, **kw)
    a3 = Arena([None, None], st, **kw)
    a1 = Arena(None, st, **kw)
    differs = False
    for _ in range(6):
        a1.step()
        a2.step()
        a3.step()
        assert [s.fen() for s in a2.states] == [s.fen() for s in a3.states]
        differs = differs or [s.fen() for s in a1.states] != [s.fen() for s in a2.states]
    assert a2.nodes == a3.nodes > 0 and differs
    # within an arena the games differ from each other too (per-tree generators): not four copies of one game
    assert len({s.fen() for s in a2.states}) > 1
    for a in (a1, a2, a3):
        a.close()
