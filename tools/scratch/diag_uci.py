import os, sys, time
sys.path.insert(0, '/root/repo')
from tests.test_uci_host_logic import Engine
for wave in ("0", "1"):
    env = dict(os.environ); env["ARA_WAVE"] = wave
    e = Engine(env)
    e.send("uci"); e.read_until("uciok", timeout=60)
    e.send("setoption name UCI_Variant value crazyhouse")
    e.send("setoption name Batch_Size value 16")
    e.send("position startpos moves e2e4")
    e.send("go infinite")
    time.sleep(0.5)
    e.send("isready"); e.read_until("readyok", timeout=60)
    e.send("stop")
    print("WAVE", wave, e.read_until("bestmove", timeout=30))
    for l in e.lines[-6:]: print("   ", l)
    e.send("quit")
