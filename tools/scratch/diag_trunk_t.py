"""bit comparison of the transposed trunk kernel against rise_trunk_kernel<64>, and timing of ara_net_predict"""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

def child(mode, outdir):
    from crazyara_b200.nn import NeuralNetAPI
    from crazyara_b200.weights import export_blob
    from crazyara_b200 import synthetic
    for name in ("risev2", "risev33"):
        arch = synthetic.risev2(34, 81) if name == "risev2" else synthetic.risev33(52, 76)
        blob = export_blob(synthetic.random_state_dict(arch, 0), arch, os.path.join(outdir, f"{name}.arab"), input_version=10 if name == "risev2" else 30)
        for B in (64, 5, 1):
            net = NeuralNetAPI("gpu", 0, B, blob)
            x = np.random.default_rng(B).random((B, arch["in_channels"], 8, 8), dtype=np.float32)
            val = np.zeros(B, np.float32)
            prob = np.zeros((B, arch["policy_channels"] * 64), np.float32)
            net.predict(x, val, prob)
            np.save(os.path.join(outdir, f"{name}_{B}_{mode}_v.npy"), val)
            np.save(os.path.join(outdir, f"{name}_{B}_{mode}_p.npy"), prob)
            if B == 64:
                for _ in range(10): net.predict(x, val, prob)
                t0 = time.perf_counter()
                for _ in range(200): net.predict(x, val, prob)
                print(f"mode {mode} {name} B={B}: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per predict (host round trip)", flush=True)

if len(sys.argv) > 2:
    child(sys.argv[1], sys.argv[2])
else:
    d = tempfile.mkdtemp()
    for mode in ("0", "1"):
        env = dict(os.environ, ARA_TRUNK_T=mode)
        r = subprocess.run([sys.executable, __file__, mode, d], env=env, capture_output=True, text=True, timeout=100)
        print(r.stdout[-2000:], r.stderr[-2000:])
    for name in ("risev2", "risev33"):
        for B in (64, 5, 1):
            for k in ("v", "p"):
                a = np.load(os.path.join(d, f"{name}_{B}_0_{k}.npy")); b = np.load(os.path.join(d, f"{name}_{B}_1_{k}.npy"))
                same = np.array_equal(a.view(np.uint32), b.view(np.uint32))
                print(name, B, k, "bit-identical" if same else f"DIFFERENT max abs {np.abs(a - b).max():.3e} nan {np.isnan(b).sum()}")
