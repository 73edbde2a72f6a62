"""The C++ host (UCI front-end over the C-ABI) drives the same device search as the Python mirror."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_uci_binary_matches_python_agent():
    exe = os.path.join(ROOT, "crazyara_b200", "ara_uci")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "crazyara_b200/ara_uci"], check=True)
    script = "\n".join(["uci", "setoption name UCI_Variant value crazyhouse", "setoption name Batch_Size value 8",
                        "setoption name Simulations value 200", "setoption name Centi_Node_Temperature value 100",
                        "isready", "position startpos moves e2e4 e7e5", "go", "root", "quit"]) + "\n"
    out = subprocess.run([exe], input=script, capture_output=True, text=True, timeout=120).stdout
    assert "uciok" in out and "readyok" in out
    best = [l for l in out.splitlines() if l.startswith("bestmove")][0].split()[1]
    from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
    st = BoardState().set("", False, 1).do_uci("e2e4", "e7e5")
    agent = MCTSAgent(None, default_settings("crazyhouse", batch_size=8, simulations=200, node_policy_temperature=1.0), 0, 1)
    r = agent.evaluate_board_state(st)
    assert r["best_move"] == best
    nodes = int([l for l in out.splitlines() if l.startswith("info depth")][0].split("nodes")[1].split()[0])
    assert nodes == r["nodes"]
