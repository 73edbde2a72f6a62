"""Raw network throughput (the reference's UCI `inference` command, engine/src/uci/crazyara.cpp:156-181)."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crazyara_b200.nn import NeuralNetAPI  # noqa: E402
from crazyara_b200.weights import export_blob  # noqa: E402
from crazyara_b200 import synthetic


def main():
    import torch
    iters = int(os.environ.get("ITERS", "200"))
    for name, arch, ver in (("risev2", synthetic.risev2(34, 81), 10), ("risev33", synthetic.risev33(52, 76, True), 30)):
        sd = synthetic.random_state_dict(arch, 0)
        with tempfile.TemporaryDirectory() as d:
            blob = export_blob(sd, arch, os.path.join(d, "w.arab"), input_version=ver)
            for batch in (1, 8, 64, 128):
                net = NeuralNetAPI("gpu", 0, batch, blob)
                C = arch["in_channels"]
                x = torch.rand(batch, C, 8, 8).pin_memory()
                v = torch.empty(batch).pin_memory()
                p = torch.empty(batch, arch["policy_channels"] * 64).pin_memory()
                xn, vn, pn = x.numpy(), v.numpy(), p.numpy()
                for _ in range(5):
                    net.predict(xn, vn, pn, None)
                t0 = time.perf_counter()
                for _ in range(iters):
                    net.predict(xn, vn, pn, None)
                t_host = (time.perf_counter() - t0) / iters
                xd = x.cuda()
                for _ in range(5):
                    net.forward_device(xd.data_ptr(), batch)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(iters):
                    net.forward_device(xd.data_ptr(), batch)
                t_dev = (time.perf_counter() - t0) / iters
                print(f"{name} B={batch:4d} host-api {t_host*1e6:8.1f} us ({batch/t_host:10.0f} evals/s)  "
                      f"device {t_dev*1e6:8.1f} us ({batch/t_dev:10.0f} evals/s)", flush=True)
                net.close()


if __name__ == "__main__":
    main()
