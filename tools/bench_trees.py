"""Aggregate search throughput with T game trees searched concurrently on one GPU (the self-play / analysis-server
shape: every tree has its own Batch_Size-64 mini-batches, one network forward serves all trees of an iteration).
TREES="1,4,16" SIMS=3200 python tools/bench_trees.py"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
from crazyara_b200.nn import NeuralNetAPI
from crazyara_b200.weights import export_blob
from crazyara_b200 import synthetic

FLOP_POS = None


def main():
    trees = [int(x) for x in os.environ.get("TREES", "1,4,16,32").split(",")]
    sims = int(os.environ.get("SIMS", "3200"))
    batch = int(os.environ.get("BATCH", "64"))
    reps = int(os.environ.get("REPS", "5"))
    arch = synthetic.risev2(34, 81)
    import bench
    flops = bench.net_flops_per_position(arch)
    d = tempfile.mkdtemp()
    blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(d, "w.arab"), input_version=10)
    openings = ["", "e2e4", "d2d4", "g1f3", "e2e4 e7e5", "d2d4 d7d5", "c2c4", "b1c3"]
    for T in trees:
        net = NeuralNetAPI("gpu", 0, batch * T, blob)
        agent = MCTSAgent(net, default_settings("crazyhouse", batch_size=batch, simulations=sims), 0, T)
        agent.set_profile(True)
        states = []
        for t in range(T):
            s = BoardState().set("", False, 1)
            mv = openings[t % len(openings)]
            if mv:
                s.do_uci(*mv.split())
            states.append(s)
        best = None
        for rep in range(reps + 2):
            for t, s in enumerate(states):
                agent.set_position(s, t)
            agent.evaluate_board_state()
            if rep < 2:
                continue
            ms = agent.last_go_ms()
            nodes = sum(r["nodes"] for r in agent.results())
            prof = agent.profile()
            row = dict(trees=T, nodes=nodes, go_ms=ms, nps=nodes / (ms * 1e-3), **prof)
            evals = sum(r["evals"] for r in agent.results())
            row["conv_tflops"] = evals * flops / (prof["net_ms"] * 1e-3) / 1e12
            if best is None or row["nps"] > best["nps"]:
                best = row
        print(json.dumps(best), flush=True)
        agent.close()
        net.close()


if __name__ == "__main__":
    main()
