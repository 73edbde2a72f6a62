"""Where a self-play step goes: device search split (select / network / apply) against the host work around it.
    python tools/prof_selfplay.py [n_games] [steps]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crazyara_b200.nn import NeuralNetAPI
from crazyara_b200.selfplay import Arena, rl_settings
from crazyara_b200.weights import export_blob
from crazyara_b200 import synthetic

n_games = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 1
arch = synthetic.risev2(34, 81)
blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(tempfile.mkdtemp(), "w.arab"), input_version=10)
st = rl_settings("crazyhouse")
nets = [NeuralNetAPI("gpu", 0, n_games // groups * st.batch_size, blob) for _ in range(groups)]
arena = Arena(nets if groups > 1 else nets[0], st, variant=1, n_games=n_games, max_plies=160, seed=1)
arena.run(max_steps=2)
if groups == 1:
    arena.agent.set_profile(True)
arena.search_ms = 0.0
t0 = time.perf_counter()
acc = dict(select_ms=0.0, net_ms=0.0, apply_ms=0.0, net_forwards=0)
for _ in range(steps):
    arena.step()
    if groups == 1:
        p = arena.agent.profile()
        for k in acc:
            acc[k] += p[k]
wall = (time.perf_counter() - t0) * 1e3
print(f"{n_games} games in {groups} group(s), {steps} steps: wall {wall / steps:.1f} ms/step, device search {arena.search_ms / steps:.1f} ms/step "
      f"(select {acc['select_ms'] / steps:.1f}, network {acc['net_ms'] / steps:.1f}, apply {acc['apply_ms'] / steps:.1f}, "
      f"{acc['net_forwards'] / steps:.0f} forwards), host {(wall - arena.search_ms) / steps:.1f} ms/step, "
      f"{n_games * steps / wall * 1e3:.0f} moves/s")
