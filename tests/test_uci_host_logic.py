"""The UCI front-end's command loop on a box without a GPU: the search entry points of the C-ABI are replaced by a
stand-in (tests/uci_stub/stub_search.c, LD_PRELOADed) that waits for its move time or a stop; everything else -- option
handling, position tracking, the worker thread of `go infinite`, `stop`, the order of the answers -- is the real binary."""
import os
import subprocess
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "crazyara_b200", "ara_uci")
STUB_SRC = os.path.join(ROOT, "tests", "uci_stub", "stub_search.c")
STUB = os.path.join(ROOT, "tests", "uci_stub", "libstub_search.so")


@pytest.fixture(scope="module")
def env():
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", ROOT, "crazyara_b200/ara_uci"], check=True)
    if not os.path.exists(STUB) or os.path.getmtime(STUB) < os.path.getmtime(STUB_SRC):
        subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-Wall", "-I" + os.path.join(ROOT, "include"), STUB_SRC, "-o", STUB],
                       check=True)
    e = dict(os.environ)
    e["LD_PRELOAD"] = STUB
    return e


class Engine:
    def __init__(self, env):
        self.p = subprocess.Popen([EXE], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1, env=env)
        self.lines = []

    def send(self, line):
        self.p.stdin.write(line + "\n")
        self.p.stdin.flush()

    def read_until(self, prefix, timeout=10.0):
        t0 = time.time()
        while time.time() - t0 < timeout:
            line = self.p.stdout.readline()
            if not line:
                break
            self.lines.append(line.rstrip("\n"))
            if self.lines[-1].startswith(prefix):
                return self.lines[-1]
        raise AssertionError(f"no line starting with {prefix!r}; got {self.lines[-8:]}")

    def close(self):
        self.p.stdin.close()
        self.p.wait(timeout=10)


def test_go_infinite_answers_isready_and_stops(env):
    e = Engine(env)
    e.send("uci")
    e.read_until("uciok")
    e.send("position startpos")
    e.send("go infinite")
    time.sleep(0.15)
    e.send("isready")
    assert e.read_until("readyok") and not any(l.startswith("bestmove") for l in e.lines)   # still searching
    t0 = time.time()
    e.send("stop")
    assert e.read_until("bestmove") == "bestmove e2e4" and time.time() - t0 < 1.0
    info = [l for l in e.lines if l.startswith("info depth")][-1]
    assert 100 <= int(info.split(" time ")[1].split()[0]) < 3000           # it ran until the stop, not to the bound
    # the engine is usable again: a synchronous search, then a second infinite one ended by `quit`
    e.send("go movetime 60")
    assert e.read_until("bestmove") == "bestmove e2e4"
    assert [l for l in e.lines if l.startswith("info string movetime")] == ["info string movetime 40"]
    e.send("go infinite")
    time.sleep(0.05)
    e.send("quit")
    assert e.read_until("bestmove") == "bestmove e2e4"
    e.close()
    assert e.p.returncode == 0


def test_commands_during_an_infinite_search_end_it_first(env):
    e = Engine(env)
    e.send("position startpos moves e2e4")
    e.send("go infinite")
    time.sleep(0.05)
    e.send("position startpos moves e2e4 e7e5")     # needs the engine: the running search is stopped and reported first
    assert e.read_until("bestmove")
    e.send("stop")                                   # nothing is running any more: ignored
    e.send("go movetime 30")
    assert e.read_until("bestmove")
    assert len([l for l in e.lines if l.startswith("bestmove")]) == 2
    e.send("quit")
    e.close()
    assert e.p.returncode == 0


def test_scripted_session_keeps_the_synchronous_order(env):
    script = "\n".join(["uci", "isready", "position startpos", "go movetime 25", "go nodes 100", "root", "quit"]) + "\n"
    out = subprocess.run([EXE], input=script, capture_output=True, text=True, timeout=30, env=env).stdout.splitlines()
    kinds = [l.split()[0] + (" " + l.split()[1] if l.startswith("info") else "") for l in out if l and not l.startswith("option")]
    assert kinds[:2] == ["id", "id"] and "uciok" in kinds and "readyok" in kinds
    tail = kinds[kinds.index("readyok") + 1:]
    assert tail[:5] == ["info string", "info depth", "bestmove", "info depth", "bestmove"]


def _session(env, tmp_path, commands):
    log = tmp_path / "calls.log"
    e = dict(env)
    e["ARA_STUB_LOG"] = str(log)
    out = subprocess.run([EXE], input="\n".join(commands + ["quit"]) + "\n", capture_output=True, text=True, timeout=60, env=e).stdout
    return out.splitlines(), (log.read_text().splitlines() if log.exists() else [])


def test_position_extension_walks_the_kept_tree(env, tmp_path):
    """Reuse_Tree: a `position` that extends the searched game announces exactly the new moves to the tree."""
    head = ["setoption name UCI_Variant value chess", "isready"]
    _, calls = _session(env, tmp_path, head + ["position startpos", "go movetime 20",
                                               "position startpos moves e2e4 e7e5", "go movetime 20",
                                               "position startpos moves d2d4", "go movetime 20",          # another game: nothing kept
                                               "ucinewgame", "position startpos moves e2e4", "go movetime 20"])
    applied = [c for c in calls if c.startswith("apply_move")]
    assert applied == ["apply_move tree=0 from=12 to=28 flag=0", "apply_move tree=0 from=52 to=36 flag=0"]   # e2e4, e7e5
    assert len([c for c in calls if c.startswith("go ")]) == 4
    (tmp_path / "calls.log").unlink()
    _, calls = _session(env, tmp_path, head[:1] + ["setoption name Reuse_Tree value false", "isready", "position startpos",
                                                   "go movetime 20", "position startpos moves e2e4 e7e5", "go movetime 20"])
    assert [c for c in calls if c.startswith("apply_move")] == []


def test_clock_games_run_under_the_time_manager(env, tmp_path):
    """wtime/btime: the move time of ara_time_for_move goes to the search AND to the ThreadManager rules; plain
    `go movetime` switches them off; the per-game NPS estimate is known from the second clock search on."""
    lines, calls = _session(env, tmp_path, ["setoption name UCI_Variant value chess", "isready", "position startpos",
                                            "go wtime 3000 btime 3000 winc 100 binc 100",
                                            "go wtime 2900 btime 3000 movestogo 40",
                                            "go movetime 50"])
    # (3000 - 20*30)/(38-1) + 0.7*100 - 20 = 64 + 70 - 20 = 114;  (2900 - 600)/40 - 20 = 37
    assert [l for l in lines if l.startswith("info string movetime")] == ["info string movetime 114", "info string movetime 37",
                                                                          "info string movetime 30"]
    tc = [c for c in calls if c.startswith("time_control")]
    assert tc == ["time_control off", "time_control movetime=114 in_game=1 can_prolong=1 nps_known=0",
                  "time_control off", "time_control movetime=37 in_game=1 can_prolong=1 nps_known=1", "time_control off"]
    assert [c for c in calls if c.startswith("go ")] == ["go movetime=114", "go movetime=37", "go movetime=30"]
