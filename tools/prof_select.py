"""Phase breakdown of the select kernel (clock64 counters inside the kernel)."""
import ctypes, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crazyara_b200 import lib
from crazyara_b200.engine import BoardState, MCTSAgent, default_settings
from crazyara_b200.nn import NeuralNetAPI
from crazyara_b200.weights import export_blob
from crazyara_b200 import synthetic

arch = synthetic.risev2(34, 81)
d = tempfile.mkdtemp()
blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(d, "w.arab"), input_version=10)
net = NeuralNetAPI("gpu", 0, 64, blob)
agent = MCTSAgent(net, default_settings("crazyhouse", batch_size=64, simulations=3200), 0, 1)
agent.set_profile(True)
for _ in range(3):
    r = agent.evaluate_board_state(BoardState().set("", False, 1))
out = (ctypes.c_ulonglong * 8)()
lib().ara_search_debug_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
lib().ara_search_debug_cycles(agent._h, 0, out)
names = ["descent", "copy+do_move", "rep+movegen", "node init", "planes", "terminal backup", "bookkeeping", "-"]
fine = "fine" in os.environ.get("ARA_B200_LIB", "")
if fine:  # -DARA_PROF_FINE build: slots 4/7/5 are sub-intervals of the descent
    names[4], names[7], names[5] = "  descent: edge wait", "  descent: puct+argmax", "  descent: child hdr wait (+term. backup)"
if os.environ.get("ARA_WAVE", "1") != "0":  # the wavefront kernel's own slots (search_wave.cuh), warp-cycles over all its warps
    fine = False
    names = ["start gate", "ply hand-off wait", "steps", "leaf preparation", "commit wait", "commit", "abort answered", "(playouts taken back)"]
tot = sum(out) - (out[4] + out[7] + out[5] if fine else 0)
print("profile", agent.profile(), "go ms", agent.last_go_ms(), "avg depth", r["sum_depth"] / max(1, r["visit_sum"]),
      "sum_k/visit", r["sum_select_k"] / max(1, r["visit_sum"]), "tree nodes", r["tree_nodes"])
for n, c in zip(names, out):
    print(f"{n:16s} {c/1e6:10.2f} Mcycles  {100.0*c/max(1,tot):5.1f}%  ({c/1.965e6:8.2f} ms at 1965 MHz)")
