"""Diagnostic: tree-stream time of a headline search with Threads 1 / 2, fake backend vs the real network, and the tower
kernel's two CTA shapes (ARA_TRUNK_ROWS) -- separates the cost of the two-thread schedule itself from the interference
of the network kernels with the select warp."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(threads, real, reps=4):
    from crazyara_b200 import synthetic
    from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
    from crazyara_b200.nn import NeuralNetAPI
    from crazyara_b200.weights import export_blob
    net = None
    if real:
        arch = synthetic.risev2(34, 81)
        blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(tempfile.mkdtemp(), "n.arab"), input_version=10)
        net = NeuralNetAPI("gpu", 0, 64, blob)
    st = default_settings("crazyhouse", batch_size=64, simulations=3200, threads=threads)
    agent = MCTSAgent(net, st, 0, 1)
    out = []
    for rep in range(reps + 2):
        agent.set_profile(rep >= reps)
        r = agent.evaluate_board_state(BoardState().set("", False, 1))
        if rep >= 1:
            prof = agent.profile() if rep >= reps else None
            out.append((agent.last_go_ms(), r["nodes"], prof))
    agent.close()
    if net:
        net.close()
    ms = min(o[0] for o in out[:reps - 1])
    prof = out[-1][2]
    print(f"threads={threads} real={int(real)} env={ {k: v for k, v in os.environ.items() if k.startswith("ARA_")} }: {ms:.2f} ms/search "
          f"({out[0][1] / ms:.1f} k NPS); profiled: tree {prof['select_ms']:.2f} net {prof['net_ms']:.2f} apply {prof['apply_ms']:.2f} ms", flush=True)


if __name__ == "__main__":
    sel = os.environ.get("DIAG", "all")
    if sel == "all":
        for threads in (1, 2):
            for real in (False, True):
                run(threads, real)
    else:
        run(2, True, reps=3)
