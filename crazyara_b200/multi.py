"""One process per GPU: how independent searches / self-play games are spread over ranks and how their counters are
combined.  The hot path has no exchange step (every search tree and every game lives on one GPU, like one engine
process of the reference per device, engine/src/rl/selfplay.cpp:225-262 runs `number of games` independent games), so
there is no data-path collective: ranks only meet at the barrier around the timed region and in the final reduction
of their counters.  Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""


def shard_range(n_items, rank, world):
    """Contiguous share of `n_items` independent units (games, positions) for `rank`: sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def aggregate_counters(nodes, dev_ms, wall_s, launches, dist=None, device="cpu"):
    """Whole-job totals of one timed region: units and launches are summed over ranks, times are the max over ranks.
    Returns (total_nodes, max_dev_ms, max_wall_s, total_launches)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(nodes), float(dev_ms), float(wall_s), int(launches)
    import torch
    s = torch.tensor([float(nodes), float(launches)], device=device, dtype=torch.float64)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    m = torch.tensor([float(dev_ms), float(wall_s)], device=device, dtype=torch.float64)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return s[0].item(), m[0].item(), m[1].item(), int(s[1].item())


def throughput(total_units, max_ms):
    return total_units / (max_ms * 1e-3) if max_ms > 0 else 0.0
