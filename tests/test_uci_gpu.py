"""The C++ host (UCI front-end over the C-ABI) drives the same device search as the Python mirror."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the UCI defaults are the reference's (Threads 2, epsilon-greedy 5 %, epsilon-checks 1 %); comparisons with a Python agent
# built from default_settings() (the deterministic single-threaded set) switch them off
PARITY = ["setoption name Threads value 1", "setoption name Centi_Epsilon_Greedy value 0", "setoption name Centi_Epsilon_Checks value 0"]


@pytest.mark.gpu
def test_uci_binary_matches_python_agent():
    exe = os.path.join(ROOT, "crazyara_b200", "ara_uci")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", ROOT, "crazyara_b200/ara_uci"], check=True)
    script = "\n".join(["uci", "setoption name UCI_Variant value crazyhouse", "setoption name Batch_Size value 8", *PARITY,
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


@pytest.mark.gpu
def test_uci_time_managed_go():
    """`go movetime` / `go wtime btime`: TimeManager's move time (overhead 20 ms subtracted) bounds the search.
    (`ucinewgame` clears the game's NPS estimate, so the clock search runs its whole move time: no early stop.)"""
    import time
    exe = os.path.join(ROOT, "crazyara_b200", "ara_uci")
    script = "\n".join(["uci", "setoption name UCI_Variant value crazyhouse", "setoption name Batch_Size value 16",
                        "setoption name Timed_Search_Nodes value 120000", "isready", "position startpos",
                        "go movetime 220", "ucinewgame", "go wtime 20000 btime 20000 winc 100 binc 100",
                        "go nodes 300", "quit"]) + "\n"
    t0 = time.time()
    out = subprocess.run([exe], input=script, capture_output=True, text=True, timeout=180).stdout
    wall = time.time() - t0
    lines = out.splitlines()
    mts = [int(l.split()[-1]) for l in lines if l.startswith("info string movetime")]
    # movetime 220 - 20; sudden death: (20000 - 600) / (38 - 1) + 0.7 * 100 - 20 = 524 + 70 - 20
    assert mts == [200, 574]
    infos = [l for l in lines if l.startswith("info depth")]
    assert len(infos) == 3 and len([l for l in lines if l.startswith("bestmove")]) == 3
    times = [int(l.split(" time ")[1].split()[0]) for l in infos]
    nodes = [int(l.split(" nodes ")[1].split()[0]) for l in infos]
    assert 150 <= times[0] <= 450 and 500 <= times[1] <= 900 and nodes[1] > nodes[0] > 1000
    assert nodes[2] >= 300 and nodes[2] < 400   # back to a visit budget: the time limit is off again
    assert wall < 60


@pytest.mark.gpu
def test_uci_position_extension_reuses_the_tree():
    """`position ... moves` that extends the searched game walks the kept tree (Reuse_Tree, apply_move_to_tree for our
    move and the reply); the second search equals the Python agent doing the same, and reports the reused nodes."""
    from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
    st = BoardState().set("", False, 1)
    agent = MCTSAgent(None, default_settings("crazyhouse", batch_size=8, simulations=300, node_policy_temperature=1.0), 0, 1, 4096)
    r0 = agent.evaluate_board_state(st)
    best, reply = r0["pv"][0], r0["pv"][1]
    for m in (best, reply):
        agent.apply_move_to_tree(m)
        st.do_uci(m)
    r1 = agent.evaluate_board_state(st)
    assert r1["nodes_pre_search"] > 0
    exe = os.path.join(ROOT, "crazyara_b200", "ara_uci")
    head = ["uci", "setoption name UCI_Variant value crazyhouse", "setoption name Batch_Size value 8", *PARITY,
            "setoption name Simulations value 300", "setoption name Centi_Node_Temperature value 100", "isready"]
    game = ["position startpos", "go", f"position startpos moves {best} {reply}", "go", "quit"]
    out = subprocess.run([exe], input="\n".join(head + game) + "\n", capture_output=True, text=True, timeout=120).stdout
    lines = out.splitlines()
    assert [l for l in lines if l.startswith("info string reused")] == [f"info string reused {r1['nodes_pre_search']} nodes"]
    bests = [l.split()[1] for l in lines if l.startswith("bestmove")]
    nodes = [int(l.split(" nodes ")[1].split()[0]) for l in lines if l.startswith("info depth")]
    assert bests == [best, r1["best_move"]] and nodes == [r0["nodes"], r1["nodes"]]
    # with the option off the second search starts from scratch
    off = head[:-1] + ["setoption name Reuse_Tree value false", "isready"] + game
    out = subprocess.run([exe], input="\n".join(off) + "\n", capture_output=True, text=True, timeout=120).stdout
    assert "info string reused" not in out
    agent.close()


@pytest.mark.gpu
def test_uci_benchmark_and_inference_commands(tmp_path):
    """`benchmark <movetime>` (crazyara.cpp:287-330: the 15 tactical positions, pass count, NPS mean / median) and
    `inference` (crazyara.cpp:156-181: evaluations per second of the loaded network)."""
    from crazyara_b200 import synthetic
    from crazyara_b200.weights import export_blob
    exe = os.path.join(ROOT, "crazyara_b200", "ara_uci")
    head = ["uci", "setoption name UCI_Variant value crazyhouse", "setoption name Batch_Size value 16",
            "setoption name Timed_Search_Nodes value 60000"]
    out = subprocess.run([exe], input="\n".join(head + ["benchmark 60", "quit"]) + "\n", capture_output=True, text=True,
                         timeout=180).stdout
    lines = out.splitlines()
    verdicts = [l for l in lines if l.startswith(("passed      -- ", "failed      -- "))]
    assert len(verdicts) == 15 and len([l for l in lines if l.startswith("alternative -- ")]) == 15
    assert any(l.startswith("Passed:") and l.endswith("/15") for l in lines)
    nps = [int(l.split()[-1]) for l in lines if l.startswith(("NPS (avg):", "NPS (median):"))]
    assert len(nps) == 2 and min(nps) > 1000
    arch = synthetic.risev2(34, 81)
    blob = export_blob(synthetic.random_state_dict(arch, 0), arch, str(tmp_path / "risev2.arab"), input_version=10)
    script = head + [f"setoption name Model_Path value {blob}", "isready", "inference warmup 2 iterations 6", "quit"]
    out = subprocess.run([exe], input="\n".join(script) + "\n", capture_output=True, text=True, timeout=180).stdout
    assert "info string batch-size: 16" in out
    rate = [float(l.split()[-2]) for l in out.splitlines() if l.startswith("info string Evaluations per second:")]
    assert len(rate) == 1 and rate[0] > 100.0


@pytest.mark.gpu
def test_uci_go_infinite_and_stop_on_the_device():
    """`go infinite` on the real device search: `isready` is answered while it runs, `stop` ends it (ara_search_stop inside
    the go loop) and the best move of the tree searched so far comes back; a `stop` during a timed `go` ends that too."""
    import time
    from tests.test_uci_host_logic import Engine
    exe = os.path.join(ROOT, "crazyara_b200", "ara_uci")
    e = Engine(dict(os.environ))
    e.send("uci")
    e.read_until("uciok", timeout=60)
    e.send("setoption name UCI_Variant value crazyhouse")
    e.send("setoption name Batch_Size value 16")
    e.send("position startpos moves e2e4")
    e.send("go movetime 100")                             # (creates the device context and the handle for time-limited searches --
    e.read_until("bestmove", timeout=120)                 #  gigabytes of node pools: seconds on a cold box -- ahead of the timed part)
    e.lines.clear()
    e.send("go infinite")
    time.sleep(0.5)
    e.send("isready")
    assert e.read_until("readyok", timeout=60) and not any(l.startswith("bestmove") for l in e.lines)
    t0 = time.time()
    e.send("stop")
    best = e.read_until("bestmove", timeout=30)
    assert time.time() - t0 < 2.0 and len(best.split()[1]) >= 4
    info = [l for l in e.lines if l.startswith("info depth")][-1]
    nodes = int(info.split(" nodes ")[1].split()[0])
    assert nodes > 1000                                   # half a second of searching on the fake backend
    # a timed search is stoppable too (every go runs on the worker thread)
    e.send("go movetime 5000")
    time.sleep(0.3)
    t0 = time.time()
    e.send("stop")
    e.read_until("bestmove", timeout=30)
    assert time.time() - t0 < 2.0
    # and a limited search that is not stopped runs to its limit while the next command waits for it
    e.send("setoption name Simulations value 400")
    e.send("go nodes 400")
    e.send("isready")
    e.read_until("bestmove", timeout=60)
    e.send("quit")
    e.close()
    assert e.p.returncode == 0 and os.path.exists(exe)
