"""N>1 path on CPU: two gloo ranks run independent oracle searches (the stand-in for one search per GPU), shard a
list of games, and combine their counters the way bench.py does (sum of units, max of times)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crazyara_b200.multi import aggregate_counters, shard_range, throughput


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_single_process_aggregate_is_identity():
    assert aggregate_counters(10, 2.0, 0.5, 7) == (10.0, 2.0, 0.5, 7)
    assert throughput(1000, 500.0) == 2000.0
    assert throughput(1000, 0.0) == 0.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import search as osearch
        from oracle.chess import Position
        # every rank owns its share of 5 "games" and searches them independently
        games = [("chess", "chess", None, []), ("crazyhouse", "crazyhouse", None, []),
                 ("chess", "chess", "r1bqkbnr/pppp1ppp/2n5/4p3/4P3/5N2/PPPP1PPP/RNBQKB1R w KQkq - 2 3", []),
                 ("kingofthehill", "lichess", None, []), ("crazyhouse", "crazyhouse", None, ["e2e4"])]
        lo, hi = shard_range(len(games), rank, world)
        nodes = 0
        for variant, mode, fen, premoves in games[lo:hi]:
            st = osearch.default_settings(mode, batch_size=8, simulations=64)
            pos = Position(fen, variant, False)
            for u in premoves:
                pos.push_uci(u)
            S = osearch.Search(st)
            r = S.run(pos, osearch.fake_net(S.n_labels), with_keys=True)
            nodes += r["nodes"]
        dev_ms = 10.0 * (rank + 1)   # deterministic stand-ins for the device / wall timers
        wall = 0.02 * (world - rank)
        dist.barrier()
        total, max_ms, max_wall, launches = aggregate_counters(nodes, dev_ms, wall, 3 + rank, dist, "cpu")
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, nodes))
        if rank == 0:
            out.put((total, max_ms, max_wall, launches, gathered))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gloo_ranks_sum_units_and_take_max_time():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    total, max_ms, max_wall, launches, gathered = out.get(timeout=10)
    assert [g[:2] for g in gathered] == [(0, 3), (3, 5)]
    assert total == float(sum(g[2] for g in gathered)) and total > 0
    assert max_ms == 20.0 and max_wall == pytest.approx(0.04)
    assert launches == 3 + 4
    assert throughput(total, max_ms) == total / 0.02
