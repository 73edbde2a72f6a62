"""Per-role cycle breakdown of the persistent trunk kernel (CTA 0), from the -DARA_TRUNK_PROF build:
    make tprof && ARA_B200_LIB=build/libara_b200_tprof.so python tools/prof_trunk.py [arch] [batch]"""
import ctypes
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crazyara_b200 import lib
from crazyara_b200.nn import NeuralNetAPI
from crazyara_b200.weights import export_blob
from crazyara_b200 import synthetic

arch_name = sys.argv[1] if len(sys.argv) > 1 else "risev2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
arch = synthetic.risev2(34, 81) if arch_name == "risev2" else synthetic.risev33(52, 76)
d = tempfile.mkdtemp()
blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(d, "w.arab"), input_version=10 if arch_name == "risev2" else 30)
net = NeuralNetAPI("gpu", 0, B, blob)
x = np.random.default_rng(0).random((B, arch["in_channels"], 8, 8), dtype=np.float32)
val = np.zeros(B, np.float32)
prob = np.zeros((B, arch["policy_channels"] * 64), np.float32)
for _ in range(5):
    net.predict(x, val, prob)
out = (ctypes.c_ulonglong * 32)()
L = lib()
L.ara_net_debug_trunk_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
if L.ara_net_debug_trunk_cycles(net._h, out) != 0:
    raise SystemExit(L.ara_last_error().decode())
mma = ["issue + misc", "wait H2 (compute warps)", "wait W2 ring", "wait X tile (block boundary)", "wait D1 buffer",
       "wait W1 ring"]
cmp_ = ["X load", "SE + tile hand-over", "wait D1 (tensor core)", "TMEM read-out", "wait chunk vectors", "barrier 1",
        "H1 write", "barrier 2", "depthwise", "wait H2 buffer", "H2 write", "wait D2", "block epilogue (D2 + b2 + X)", "SE of the next block (rest)", "tile store + hand-over", "  SE: pooling", "  SE: fc1", "  SE: hidden", "  SE: fc2"]
if os.environ.get("ARA_TRUNK_T", "1") != "0" and B <= 148:  # rise_trunk_t.cuh's slots
    mma = ["issue + misc", "wait X tile (block boundary)", "wait H2 (compute warps)", "wait weight stream", "wait D1 buffer"]
    cmp_ = ["X load", "SE + tile hand-over", "wait D1 (tensor core)", "TMEM read-out + relu", "wait H2 buffer", "depthwise + H2 write",
            "wait D2", "block epilogue (D2 + b2 + X)", "SE of the next block (rest)", "tile store + hand-over", "  SE: pooling", "  SE: fc1", "  SE: hidden", "  SE: fc2"]
for title, names, base in (("MMA issuer warp", mma, 0), ("compute warp 2", cmp_, 16)):
    vals = [out[base + i] for i in range(len(names))]
    tot = sum(vals)
    print(f"{title}: {tot / 1e3:.1f} kcycles = {tot / 1.965e3:.1f} us at 1965 MHz")
    for n, v in zip(names, vals):
        print(f"    {n:34s} {v / 1e3:9.1f} kcycles {100.0 * v / max(1, tot):5.1f}%")
