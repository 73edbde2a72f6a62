#!/usr/bin/env python
"""bench.py -- headline benchmark: MCTS simulations/sec (NPS), crazyhouse start position, Batch_Size 64.

One "step" = one complete search (`go`) of --sims simulations from the crazyhouse start position with a fresh tree:
root evaluation, then ceil(sims/64) mini-batch iterations of select -> RISEv2 conv stack (tcgen05) -> scatter/backup,
all device-resident.  NPS is computed exactly like the reference: (root.visitSum - root.freeVisits) / elapsed
(engine/src/evalinfo.cpp:73-85, node.cpp:1303-1306).

  value  : device-resident NPS -- CUDA events on the search stream around each go (root board already uploaded)
  e2e    : the same searches through the public host API (BoardState -> MCTSAgent.evaluate_board_state -> EvalInfo),
           wall clock, host<->device copies inside
  roofline: conv stack (the dominant kernels): algorithmic FLOPs of the network forwards of a search / their device
           time (CUDA events around every forward on the search stream, taken live on three extra searches right after
           the timed ones -- the timed searches launch each iteration as one graph, without events in between) vs the
           measured sustained bf16 tensor peak of MEASURED_PEAKS.json
  cpu_baseline: the CPU oracle search (oracle/mcts.c, 1 thread, the reference's cost structure) with the fp32 torch
           CPU network on all host cores, on a bounded sample of the same workload (rank 0, N = 1 only)

--impl reference times that CPU arm alone (the reference engine cannot be compiled here: its move generator and
vector library are un-vendored submodules -- SURVEY 0.3 -- so the oracle port is the reference arm).
Multi-GPU: replicas only (games/searches never interact; no collective on the data path): every rank runs the same
workload on its own GPU, value = sum of nodes / max over ranks of the time ("weak" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MCTS simulations/sec (NPS) crazyhouse startpos batch=64"
UNIT = "nodes/s"


def net_flops_per_position(arch):
    """2*MAC per evaluated position (BN folded), SURVEY Appendix B."""
    C = arch["channels"]
    f = 2 * 64 * C * arch["in_channels"] * 9
    for k, se, cop in zip(arch["kernels"], arch["se_types"], arch["c_ops"]):
        f += 2 * 64 * (C * cop) * 2 + 2 * 64 * cop * k * k
        if se == "ca_se":
            f += 2 * (C * (C // 2)) * 2
        elif se == "eca_se":
            f += 2 * C * C
    f += 2 * 64 * C * C * 9 + 2 * 64 * C * arch["policy_channels"] * 9
    f += 2 * 64 * C * 8 + (2 * 512 * 4 if arch["wdl"] else 2 * (512 * 256 + 256))
    return f


class ClockSampler(threading.Thread):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.device)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_arm(sims, batch, steps, warmup):
    """Reference arm / cpu_baseline: oracle search + fp32 torch CPU network, bounded sample per step."""
    import numpy as np
    import torch

    from oracle import net as onet
    from oracle import search as osr
    from oracle.chess import Position
    # all host threads torch can use profitably: beyond ~16 threads the 8x8-board convolutions only get slower
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    try:  # cgroup CPU quota (the container may see every host core but only be allowed a few)
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            avail = min(avail, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    cores = max(1, min(avail, int(os.environ.get("ARA_CPU_THREADS", "16"))))
    torch.set_num_threads(cores)
    from crazyara_b200 import synthetic
    arch = synthetic.risev2(34, 81)              # the same random network the GPU arm runs
    sd = synthetic.random_state_dict(arch, 0)
    st = osr.default_settings("crazyhouse", batch_size=batch, simulations=sims)

    def net_fn(planes):
        out = onet.forward(sd, arch, planes)
        return out["value"], out["prob"]

    def one():
        S = osr.Search(st)
        t0 = time.perf_counter()
        r = S.run(Position(variant="crazyhouse"), net_fn)
        dt = time.perf_counter() - t0
        S.close()
        return r["nodes"], dt

    for _ in range(warmup):
        one()
    nodes, secs = 0, 0.0
    for _ in range(steps):
        n, dt = one()
        nodes += n
        secs += dt
    return nodes / secs, secs / steps * 1e3, cores, nodes


def multi_tree_leg(blob, device, trees, batch, sims, flops_pos, reps=3):
    """T independent searches (each Batch_Size `batch`) advanced together on one GPU: the analysis-server / arena shape.
    Every iteration one network forward serves all trees, so the conv stack sees T*batch positions."""
    from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
    from crazyara_b200.nn import NeuralNetAPI
    net = NeuralNetAPI("gpu", device, batch * trees, blob)
    agent = MCTSAgent(net, default_settings("crazyhouse", batch_size=batch, simulations=sims), device, trees)
    openings = ["", "e2e4", "d2d4", "g1f3", "e2e4 e7e5", "d2d4 d7d5", "c2c4", "b1c3"]
    states = []
    for t in range(trees):
        s = BoardState().set("", False, 1)
        if openings[t % len(openings)]:
            s.do_uci(*openings[t % len(openings)].split())
        states.append(s)
    best = None
    for rep in range(reps + 2):
        profiled = rep == reps + 1  # the last repetition carries events between the kernels for the phase split
        agent.set_profile(profiled)
        for t, s in enumerate(states):
            agent.set_position(s, t)
        agent.evaluate_board_state()
        if rep == 0:
            continue
        res = agent.results()
        if profiled:
            prof = agent.profile()
            evals = sum(r["evals"] for r in res)
            best.update({"net_ms": prof["net_ms"], "select_ms": prof["select_ms"], "apply_ms": prof["apply_ms"],
                         "conv_tflops": evals * flops_pos / (prof["net_ms"] * 1e-3) / 1e12})
            continue
        ms = agent.last_go_ms()
        nodes = sum(r["nodes"] for r in res)
        row = {"trees": trees, "batch_per_tree": batch, "simulations": sims, "nps": nodes / (ms * 1e-3), "ms_per_go": ms}
        if best is None or row["nps"] > best["nps"]:
            best = row
    agent.close()
    net.close()
    return best


def selfplay_leg(blob, device, n_games, seconds):
    """Self-play games/hour (second half of BASELINE.json's metric): `n_games` concurrent crazyhouse games per GPU with
    the reference's RL search settings (rl_config.py:34-65: 800 nodes, Batch_Size 8, Dirichlet 0.25/0.3)."""
    from crazyara_b200.nn import NeuralNetAPI
    from crazyara_b200.selfplay import Arena, rl_settings
    st = rl_settings("crazyhouse")
    # two groups of games, each with its own agent and network buffers, searched from two host threads: one group's
    # tree kernels overlap the other's network forward
    groups = 2 if n_games % 2 == 0 else 1
    nets = [NeuralNetAPI("gpu", device, n_games // groups * st.batch_size, blob) for _ in range(groups)]
    arena = Arena(nets, st, variant=1, n_games=n_games, device=device, max_plies=160, seed=1)
    arena.run(max_steps=2)  # warm-up (graph capture, allocations)
    arena.finished.clear()
    arena.nodes, arena.search_ms = 0, 0.0
    res = arena.run(max_seconds=seconds)
    arena.close()
    for net in nets:
        net.close()
    # random weights do not finish games the way a trained network does, so the rate is quoted per searched move and
    # converted with a nominal 100-ply game; the games that did finish inside the window are reported beside it
    return {"concurrent_games": n_games, "game_groups": groups, "settings": "RL defaults: nodes 800, Batch_Size 8, Dirichlet eps 0.25 alpha 0.3, "
            "temperature 0.8 for 15 plies, games adjudicated at 160 plies (random weights)",
            "moves_per_s": res["moves_per_s"], "games_per_hour_at_100_plies": res["moves_per_s"] * 36.0,
            "games_finished_in_window": res["games"], "avg_plies_finished": res["avg_plies"],
            "search_nps": res["nps"], "wall_s": res["wall_s"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sims", type=int, default=3200)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--threads", type=int, default=1, choices=[1, 2], help="Threads: 1 = deterministic parity mode, 2 = the "
                    "reference's default (two logical search threads: one selects while the other's batch is evaluated)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU arm (profiling runs)")
    ap.add_argument("--cpu-sims", type=int, default=1280, help="bounded CPU sample: simulations per CPU search")
    ap.add_argument("--trees", type=int, default=32, help="extra leg: concurrent searches per GPU (0 = skip)")
    ap.add_argument("--selfplay-seconds", type=float, default=8.0, help="extra leg: self-play arena window (0 = skip)")
    ap.add_argument("--selfplay-games", type=int, default=64)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    workload = (f"crazyhouse startpos, RISEv2-mobile 34x8x8 -> 81x64 policy map, Batch_Size {args.batch}, "
                f"Simulations {args.sims}, Threads 1, reference UCI defaults (node temperature 1.7, virtual_mix, "
                f"MCTS solver on, no Dirichlet/epsilon), fresh tree per step")

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 8))
        nps, ms, cores, _ = cpu_arm(args.cpu_sims, args.batch, steps, min(args.warmup, 1))
        sample = (f"{steps} searches of {args.cpu_sims} simulations (Batch_Size {args.batch}) of the same workload; "
                  f"C oracle search on 1 thread + fp32 torch CPU network on {cores} threads")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": nps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded random RISEv2 weights, start position)",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": nps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": nps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at VERSION level; stdout carries exactly one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
    from crazyara_b200.nn import NeuralNetAPI
    from crazyara_b200.weights import export_blob
    from crazyara_b200 import synthetic  # seeded random weights (no trained weights ship with the reference)

    arch = synthetic.risev2(34, 81)
    flops_pos = net_flops_per_position(arch)
    tmp = tempfile.mkdtemp(prefix="ara_bench_")
    blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(tmp, f"risev2_{rank}.arab"), input_version=10)
    net = NeuralNetAPI("gpu", local_rank, args.batch, blob)
    settings = default_settings("crazyhouse", batch_size=args.batch, simulations=args.sims, threads=args.threads)
    agent = MCTSAgent(net, settings, local_rank, 1)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step():
        flush.fill_(1)  # L2 flush between steps (outside the timed region)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        state = BoardState().set("", False, 1)
        r = agent.evaluate_board_state(state)
        wall = time.perf_counter() - t0
        return r, wall, agent.last_go_ms()

    for _ in range(args.warmup):
        step()
    launches0 = agent.launch_count() + net.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    nodes = 0
    dev_ms = wall_s = net_ms = sel_ms = app_ms = 0.0
    forwards = 0
    last = None
    for _ in range(args.steps):
        r, wall, ms = step()
        nodes += int(r["nodes"])
        dev_ms += ms
        wall_s += wall
        last = r
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = agent.launch_count() + net.launch_count() - launches0
    # phase split (select / network / apply): CUDA events between the kernels of every iteration, which the timed
    # searches above do without (an iteration is one graph launch there) -- measured on extra searches, scaled to K
    agent.set_profile(True)
    n_prof = 3
    for _ in range(n_prof):
        step()
        prof = agent.profile()
        net_ms += prof["net_ms"] * args.steps / n_prof
        sel_ms += prof["select_ms"] * args.steps / n_prof
        app_ms += prof["apply_ms"] * args.steps / n_prof
        forwards += prof["net_forwards"] * args.steps / n_prof
    agent.set_profile(False)

    from crazyara_b200.multi import aggregate_counters
    total_nodes, max_dev_ms, max_wall, launches = aggregate_counters(nodes, dev_ms, wall_s, launches, dist, "cuda")

    # secondary legs (outside the timed region of the headline number): many searches per GPU, and self-play
    agent.close()
    net.close()
    extra = {}
    if args.trees > 0:
        extra["multi_tree"] = multi_tree_leg(blob, local_rank, args.trees, args.batch, args.sims, flops_pos)
    if args.selfplay_seconds > 0:
        extra["selfplay"] = selfplay_leg(blob, local_rank, args.selfplay_games, args.selfplay_seconds)
    if dist is not None:  # whole-job figures: sums over ranks (independent replicas)
        sums = torch.tensor([extra.get("multi_tree", {}).get("nps", 0.0), extra.get("selfplay", {}).get("moves_per_s", 0.0)],
                            device="cuda", dtype=torch.float64)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        if "multi_tree" in extra:
            extra["multi_tree"]["nps_all_gpus"] = sums[0].item()
        if "selfplay" in extra:
            extra["selfplay"]["moves_per_s_all_gpus"] = sums[1].item()
            extra["selfplay"]["games_per_hour_at_100_plies_all_gpus"] = sums[1].item() * 36.0

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        conv_tflops = forwards * args.batch * flops_pos / (net_ms * 1e-3) / 1e12 if net_ms > 0 else 0.0
        traffic = None  # DRAM bytes per launch of the dominant tensor kernel, from the committed `ncu --set full` capture
        try:
            k = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_rise_trunk_kernel.json")))["kernels"][0]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = (k["dram__bytes_read.sum"] * scale[k["dram__bytes_read.sum unit"]] +
                       k["dram__bytes_write.sum"] * scale[k["dram__bytes_write.sum unit"]])
        except Exception:
            pass
        value = total_nodes / (max_dev_ms * 1e-3)
        e2e_value = total_nodes / max_wall
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": max_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 tensor-core operands, f32 accumulate (reference default Precision float16); f32/f64 search arithmetic",
            "data": "synthetic (seeded random RISEv2 weights; crazyhouse start position)",
            "config": {"workload": workload, "parallelism": f"replicas x{world} (one search per GPU, no collective)",
                       "l2_flush_between_steps": True, "nodes_per_step": nodes / args.steps,
                       "select_ms_per_step": sel_ms / args.steps, "net_ms_per_step": net_ms / args.steps,
                       "apply_ms_per_step": app_ms / args.steps, "net_forwards_per_step": forwards / args.steps,
                       "best_move": last.get("best_move"), "evals_per_step": int(last["evals"])},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": 128 + 136 + 16, "d2h_bytes_per_step": 14392 + 4 * (2 + int(last["iterations"]) // 2)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": conv_tflops / peak_tf if peak_tf else None, "traffic": traffic,
                         "traffic_note": "rise_trunk_kernel, one launch of 64 positions, dram__bytes_read+write "
                                         "(profiles/r01_ncu_rise_trunk_kernel.json; cold L2: the 6.9 MB of weights + the input tile)",
                         "kernel": "RISEv2 conv stack per forward of 64 positions: rise_trunk_kernel (13 bottleneck blocks, "
                                   "tcgen05 TS/SS MMAs, one launch) + stem/policy conv_gemm_kernel + head kernels",
                         "flop_per_position": flops_pos, "peak_source": peak_src},
        }
        try:
            # the other big kernel, against ITS roofline: select reads 32 B of header + 13 B per open child
            # (Q, n, P, vl) at every tree level -- a dependent pointer chase, so far below the HBM peak by nature
            sel_bytes = 32.0 * float(last.get("sum_depth", 0)) + 13.0 * float(last.get("sum_select_k", 0))
            hbm_peak = float(peaks.get("hbm_gbs", 6500.0))
            sel_gbs = sel_bytes / (sel_ms / args.steps * 1e-3) / 1e9 if sel_ms > 0 else 0.0
            out["roofline_select"] = {"bound": "hbm", "achieved": sel_gbs, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": sel_gbs / hbm_peak if hbm_peak else None,
                                      "algorithmic_bytes_per_search": sel_bytes,
                                      "note": "select_kernel: one warp per tree, one dependent L2/HBM round trip per tree "
                                              "level; latency-bound (profiles/r01_ncu_select_kernel.json)"}
        except Exception:
            pass
        if world == 1 and not args.no_cpu_baseline:
            nps, ms, cores, _ = cpu_arm(args.cpu_sims, args.batch, 2, 1)
            out["cpu_baseline"] = {"value": nps, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"2 searches of {args.cpu_sims} simulations (Batch_Size {args.batch}); C oracle "
                                             f"search 1 thread + fp32 torch CPU network {cores} threads"}
        if "multi_tree" in extra:
            mt = extra["multi_tree"]
            mt["conv_frac_of_peak"] = mt["conv_tflops"] / peak_tf if peak_tf else None
            out["multi_tree"] = mt
        if "selfplay" in extra:
            out["selfplay"] = extra["selfplay"]
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
