"""Weight import: reference state_dict (PyTorch parameter names of RiseV3) -> ARAB2001 blob for ara_net_create.

BatchNorm (eval mode, eps 1e-5) is folded into the preceding convolution here, in float64, so the CUDA side only
sees conv weight + bias.  The blob is architecture-described by its header, so any RISEv2 / RISEv3.x checkpoint of
the reference trainer (trainer_agent_pytorch.py:506-516 saves {'model_state_dict': ...}) converts without code
changes.  Tensor order must match crazyara_b200/csrc/net.cu (Net::init).
"""
import struct

import numpy as np

BN_EPS = 1e-5
SE_CODE = {None: 0, "ca_se": 1, "se": 1, "eca_se": 2}


def _np(t):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float64)


def _fold(sd, conv_key, bn_prefix):
    w = _np(sd[conv_key])
    g, b = _np(sd[bn_prefix + ".weight"]), _np(sd[bn_prefix + ".bias"])
    m, v = _np(sd[bn_prefix + ".running_mean"]), _np(sd[bn_prefix + ".running_var"])
    s = g / np.sqrt(v + BN_EPS)
    return w * s.reshape(-1, *([1] * (w.ndim - 1))), b - m * s


def export_blob(sd, arch, path, input_version=10):
    """arch: dict(in_channels, policy_channels, kernels[], se_types[], c_ops[], wdl) as in oracle-free product use;
    see crazyara_b200.nn.ARCH_RISEV2 / ARCH_RISEV33 helpers."""
    tensors = []

    def put(a):
        tensors.append(np.ascontiguousarray(a, dtype=np.float32).reshape(-1))

    w, b = _fold(sd, "body_spatial.0.body.0.weight", "body_spatial.0.body.1")
    put(w), put(b)
    for i, (k, se, cop) in enumerate(zip(arch["kernels"], arch["se_types"], arch["c_ops"])):
        p = f"body_spatial.{i + 1}"
        code = SE_CODE[se]
        if code == 1:
            put(_np(sd[p + ".se.fc.0.weight"])), put(_np(sd[p + ".se.fc.2.weight"]))
        elif code == 2:
            wc = _np(sd[p + ".se.body.0.weight"])
            put(wc[:, :, wc.shape[2] // 2]), put(_np(sd[p + ".se.body.0.bias"]))
        w, b = _fold(sd, p + ".body.0.weight", p + ".body.1")
        assert w.shape[0] == cop
        put(w), put(b)
        w, b = _fold(sd, p + ".body.3.weight", p + ".body.4")
        assert w.shape == (cop, 1, k, k)
        put(w), put(b)
        w, b = _fold(sd, p + ".body.6.weight", p + ".body.7")
        put(w), put(b)
    w, b = _fold(sd, "value_head.body.0.weight", "value_head.body.1")
    put(w), put(b)
    if arch["wdl"]:
        put(_np(sd["value_head.body_wdl.0.weight"])), put(_np(sd["value_head.body_wdl.0.bias"]))
        put(_np(sd["value_head.body_plys.0.weight"])), put(_np(sd["value_head.body_plys.0.bias"]))
    else:
        put(_np(sd["value_head.body_final.0.weight"])), put(_np(sd["value_head.body_final.0.bias"]))
        put(_np(sd["value_head.body_final.2.weight"])), put(_np(sd["value_head.body_final.2.bias"]))
    w, b = _fold(sd, "policy_head.body.0.weight", "policy_head.body.1")
    put(w), put(b)
    put(_np(sd["policy_head.body.3.weight"]))

    with open(path, "wb") as f:
        f.write(b"ARAB2001")
        f.write(struct.pack("<8i", arch["in_channels"], arch["policy_channels"], len(arch["kernels"]), 256, 8, 256,
                            1 if arch["wdl"] else 0, input_version))
        for k, se, cop in zip(arch["kernels"], arch["se_types"], arch["c_ops"]):
            f.write(struct.pack("<3i", cop, k, SE_CODE[se]))
        for t in tensors:
            f.write(struct.pack("<q", t.size))
            f.write(t.tobytes())
    return path


def arch_from_state_dict(sd):
    """Reads the architecture off a RiseV3 state_dict (rise_mobile_v3.py:81-183): block count, operating channels,
    depthwise kernel sizes, SE flavour per block, WDL head, input / policy channels."""
    keys = set(sd.keys())
    n_blocks = 0
    while f"body_spatial.{n_blocks + 1}.body.0.weight" in keys:
        n_blocks += 1
    if n_blocks == 0:
        raise ValueError("not a RiseV3 state_dict (no body_spatial.1.body.0.weight)")
    kernels, se_types, c_ops = [], [], []
    for i in range(1, n_blocks + 1):
        p = f"body_spatial.{i}"
        dw = sd[p + ".body.3.weight"]
        c_ops.append(int(dw.shape[0]))
        kernels.append(int(dw.shape[-1]))
        if p + ".se.fc.0.weight" in keys:
            se_types.append("ca_se")
        elif p + ".se.body.0.weight" in keys:
            se_types.append("eca_se")
        else:
            se_types.append(None)
    stem = sd["body_spatial.0.body.0.weight"]
    return dict(name=f"rise_{n_blocks}b", in_channels=int(stem.shape[1]), policy_channels=int(sd["policy_head.body.3.weight"].shape[0]),
                channels=int(stem.shape[0]), kernels=kernels, se_types=se_types, c_ops=c_ops,
                wdl="value_head.body_wdl.0.weight" in keys, value_channels=int(sd["value_head.body.0.weight"].shape[0]),
                value_fc=256)


def import_checkpoint(checkpoint_path, blob_path, input_version=None):
    """Reference trainer checkpoint (`torch.save({'model_state_dict': ...})`, trainer_agent_pytorch.py:506-516; a bare
    state_dict is accepted too) -> ARAB2001 blob.  input_version defaults from the input channel count
    (34/63 -> 1.0, 51 -> 2.0, 52/64/80 -> 3.0)."""
    import torch
    ck = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    sd = ck.get("model_state_dict", ck) if isinstance(ck, dict) else ck
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}  # DataParallel prefix
    arch = arch_from_state_dict(sd)
    if input_version is None:
        input_version = {34: 10, 63: 10, 39: 10, 51: 20, 52: 30, 64: 30, 80: 30}.get(arch["in_channels"], 10)
    export_blob(sd, arch, blob_path, input_version=input_version)
    return arch


if __name__ == "__main__":
    import sys
    if len(sys.argv) < 3:
        raise SystemExit("usage: python -m crazyara_b200.weights <checkpoint.tar|state_dict.pt> <out.arab> [input_version]")
    a = import_checkpoint(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    print(f"{sys.argv[2]}: {len(a['kernels'])} blocks, {a['in_channels']} -> {a['policy_channels']}x64, wdl={a['wdl']}")
