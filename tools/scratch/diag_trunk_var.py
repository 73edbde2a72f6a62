"""runs the bit comparison (old M=64 kernel vs the transposed kernel) for several builds of the library"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
def child(tag, outdir):
    from crazyara_b200.nn import NeuralNetAPI
    from crazyara_b200.weights import export_blob
    from crazyara_b200 import synthetic
    arch = synthetic.risev2(34, 81)
    blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(outdir, f"w{tag}.arab"), input_version=10)
    B = 64
    net = NeuralNetAPI("gpu", 0, B, blob)
    x = np.random.default_rng(B).random((B, arch["in_channels"], 8, 8), dtype=np.float32)
    val = np.zeros(B, np.float32); prob = np.zeros((B, arch["policy_channels"] * 64), np.float32)
    for rep in range(3):
        net.predict(x, val, prob)
        np.save(os.path.join(outdir, f"{tag}_{rep}_v.npy"), val); np.save(os.path.join(outdir, f"{tag}_{rep}_p.npy"), prob)
if len(sys.argv) > 2:
    child(sys.argv[1], sys.argv[2])
else:
    d = tempfile.mkdtemp()
    runs = [("old", None, "0")] + [(v, v, "1") for v in ("default", "p1", "p2", "xo")]
    for tag, lib, mode in runs:
        env = dict(os.environ, ARA_TRUNK_T=mode)
        if lib and lib != "default": env["ARA_B200_LIB"] = os.path.join(ROOT, "build", f"libara_b200_{lib}.so")
        r = subprocess.run([sys.executable, __file__, tag, d], env=env, capture_output=True, text=True, timeout=300)
        if r.returncode: print(tag, "FAILED", r.stderr[-500:])
    ref = np.load(os.path.join(d, "old_0_p.npy"))
    for tag, _, _ in runs:
        for rep in range(3):
            p = np.load(os.path.join(d, f"{tag}_{rep}_p.npy"))
            bad = [i for i in range(64) if not np.array_equal(p[i].view(np.uint32), ref[i].view(np.uint32))]
            print(tag, rep, "boards differing from the old kernel:", len(bad), bad[:20])
