"""CPU restatement (torch fp32 functional ops) of the reference's RISE network -- TEST INFRASTRUCTURE ONLY.

Follows, line by line, QueensGambit/CrazyAra
  DeepCrazyhouse/src/domain/neural_net/architectures/pytorch/rise_mobile_v3.py:81-174 (RiseV3),
    :186-214 (get_rise_v33_model), :217-241 (get_rise_v2_model), :36-78 (_get_res_blocks)
  DeepCrazyhouse/src/domain/neural_net/architectures/pytorch/builder_util.py:154-178 (_Stem),
    :437-475 (_BottlekneckResidualBlock), :83-114 (_ChannelAttentionModule), :49-80
    (_EfficientChannelAttentionModule), :206-243 (_PolicyHead), :246-326 (_ValueHead), :385-399
  and appends the softmax the engine's backend adds (engine/src/nn/tensorrtapi.cpp:378-380).

Parity pin: tests/test_oracle_net.py loads the same state_dict into the *real* reference module (importable in
the build container with a timm.DropPath shim) and compares outputs; tests/golden/net_*.json holds outputs of
the reference module for the seeded weights so the pin also holds on the GPU box where /root/reference is absent.
State dicts use the reference's own parameter names, so trained checkpoints of the reference load unchanged.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def arch_risev2(in_channels=34, policy_channels=81):
    """get_rise_v2_model (rise_mobile_v3.py:217-241): 13 blocks, k=3, c_op=128+64 i, ca_se on blocks 8..12."""
    kernels = [3] * 13
    se = [None] * 13
    for i in (8, 9, 10, 11, 12):
        se[i] = "ca_se"
    c_ops = [128 + 64 * i for i in range(13)]
    return dict(name="risev2", in_channels=in_channels, policy_channels=policy_channels, channels=256,
                kernels=kernels, se_types=se, c_ops=c_ops, wdl=False, value_channels=8, value_fc=256)


def arch_risev33(in_channels=52, policy_channels=76, wdl=True):
    """get_rise_v33_model (rise_mobile_v3.py:186-214) + _get_res_blocks channel rule (:43-49)."""
    kernels = [3] * 15
    for i in (7, 11, 12, 13):
        kernels[i] = 5
    se = [None] * 15
    for i in (5, 8, 12, 13, 14):
        se[i] = "eca_se"
    c_ops = []
    c = 224
    for i, k in enumerate(kernels):
        c_ops.append(c - 32 * (i // 2) if k == 5 else c)
        c += 32
    return dict(name="risev33", in_channels=in_channels, policy_channels=policy_channels, channels=256,
                kernels=kernels, se_types=se, c_ops=c_ops, wdl=wdl, value_channels=8, value_fc=256)


def make_state_dict(arch, seed=0):
    """Seeded random parameters under the reference's state_dict key names (no trained weights ship with the
    reference).  He-style conv init keeps activations O(1); BatchNorm statistics are perturbed so that BN folding is
    exercised."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, cout, cin, k, groups=1):
        fan_in = (cin // groups) * k * k
        sd[name] = (rng.standard_normal((cout, cin // groups, k, k)) * np.sqrt(2.0 / fan_in)).astype(np.float32)

    def bn(prefix, c):
        sd[prefix + ".weight"] = rng.uniform(0.8, 1.2, c).astype(np.float32)
        sd[prefix + ".bias"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        sd[prefix + ".running_mean"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)

    def linear(prefix, cout, cin, bias=True, scale=1.0):
        sd[prefix + ".weight"] = (rng.standard_normal((cout, cin)) * scale / np.sqrt(cin)).astype(np.float32)
        if bias:
            sd[prefix + ".bias"] = (rng.standard_normal(cout) * 0.1).astype(np.float32)

    C = arch["channels"]
    conv("body_spatial.0.body.0.weight", C, arch["in_channels"], 3)
    bn("body_spatial.0.body.1", C)
    for i, (k, se, cop) in enumerate(zip(arch["kernels"], arch["se_types"], arch["c_ops"])):
        p = f"body_spatial.{i + 1}"
        if se == "ca_se":
            linear(p + ".se.fc.0", C // 2, C, bias=False, scale=2.0)
            linear(p + ".se.fc.2", C, C // 2, bias=False, scale=2.0)
        elif se == "eca_se":
            sd[p + ".se.body.0.weight"] = (rng.standard_normal((C, C, 5)) * 2.0 / np.sqrt(C)).astype(np.float32)
            sd[p + ".se.body.0.bias"] = (rng.standard_normal(C) * 0.5).astype(np.float32)
        conv(p + ".body.0.weight", cop, C, 1)
        bn(p + ".body.1", cop)
        conv(p + ".body.3.weight", cop, cop, k, groups=cop)
        bn(p + ".body.4", cop)
        # the residual branch output is scaled down so the 13-15 block trunk stays O(1)
        sd[p + ".body.6.weight"] = (rng.standard_normal((C, cop, 1, 1)) * 0.5 * np.sqrt(1.0 / cop)).astype(np.float32)
        bn(p + ".body.7", C)
    conv("value_head.body.0.weight", arch["value_channels"], C, 1)
    bn("value_head.body.1", arch["value_channels"])
    nflat = arch["value_channels"] * 64
    if arch["wdl"]:
        linear("value_head.body_wdl.0", 3, nflat)
        linear("value_head.body_plys.0", 1, nflat)
    else:
        linear("value_head.body_final.0", arch["value_fc"], nflat, scale=1.4)
        linear("value_head.body_final.2", 1, arch["value_fc"], scale=1.4)
    conv("policy_head.body.0.weight", C, C, 3)
    bn("policy_head.body.1", C)
    sd["policy_head.body.3.weight"] = (rng.standard_normal((arch["policy_channels"], C, 3, 3)) *
                                       2.0 * np.sqrt(2.0 / (C * 9))).astype(np.float32)
    return sd


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=BN_EPS)


def forward(sd_np, arch, x):
    """x: [B, C, 8, 8] fp32 tensor/array -> dict(value [B], policy_logits [B, P*64], prob [B, P*64], aux [B,4]|None)."""
    sd = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in sd_np.items()}
    x = torch.as_tensor(x, dtype=torch.float32)
    with torch.no_grad():
        # _Stem: conv3x3 + BN + ReLU (builder_util.py:167-171)
        out = F.relu(_bn(F.conv2d(x, sd["body_spatial.0.body.0.weight"], padding=1), sd, "body_spatial.0.body.1"))
        for i, (k, se) in enumerate(zip(arch["kernels"], arch["se_types"])):
            p = f"body_spatial.{i + 1}"
            if se == "ca_se":  # builder_util.py:105-114, hard sigmoid (use_hard_sigmoid=True passed at :452)
                y = out.mean(dim=(2, 3))
                y = F.relu(F.linear(y, sd[p + ".se.fc.0.weight"]))
                y = F.hardsigmoid(F.linear(y, sd[p + ".se.fc.2.weight"]))
                out = out * y[:, :, None, None]
            elif se == "eca_se":  # builder_util.py:71-80: Conv1d(C, C, k=5, pad=2) on a length-1 sequence
                y = out.mean(dim=(2, 3))[:, :, None]
                y = F.conv1d(y, sd[p + ".se.body.0.weight"], sd[p + ".se.body.0.bias"], padding=2)
                out = out * F.hardsigmoid(y)[:, :, :, None]
            # body (builder_util.py:458-465); shortcut is the SE-scaled tensor (:473-475)
            h = F.relu(_bn(F.conv2d(out, sd[p + ".body.0.weight"]), sd, p + ".body.1"))
            h = F.relu(_bn(F.conv2d(h, sd[p + ".body.3.weight"], padding=k // 2, groups=h.shape[1]), sd, p + ".body.4"))
            h = _bn(F.conv2d(h, sd[p + ".body.6.weight"]), sd, p + ".body.7")
            out = out + h
        # _ValueHead (builder_util.py:268-326)
        v = F.relu(_bn(F.conv2d(out, sd["value_head.body.0.weight"]), sd, "value_head.body.1")).reshape(x.shape[0], -1)
        aux = None
        if arch["wdl"]:
            wdl = F.linear(v, sd["value_head.body_wdl.0.weight"], sd["value_head.body_wdl.0.bias"])
            plys = torch.sigmoid(F.linear(v, sd["value_head.body_plys.0.weight"], sd["value_head.body_plys.0.bias"]))
            sm = torch.softmax(wdl, dim=1)
            value = -sm[:, 0] + sm[:, 2]
            aux = torch.cat((wdl, plys), dim=1)
        else:
            h = F.relu(F.linear(v, sd["value_head.body_final.0.weight"], sd["value_head.body_final.0.bias"]))
            value = torch.tanh(F.linear(h, sd["value_head.body_final.2.weight"], sd["value_head.body_final.2.bias"]))[:, 0]
        # _PolicyHead, select_policy_from_plane=True (builder_util.py:225-229, :237-238)
        ph = F.relu(_bn(F.conv2d(out, sd["policy_head.body.0.weight"], padding=1), sd, "policy_head.body.1"))
        logits = F.conv2d(ph, sd["policy_head.body.3.weight"], padding=1).reshape(x.shape[0], -1)
        prob = torch.softmax(logits, dim=1)  # appended by the engine backend (tensorrtapi.cpp:378-380)
    return dict(value=value.numpy(), policy_logits=logits.numpy(), prob=prob.numpy(),
                aux=None if aux is None else aux.numpy(), trunk=out.numpy())
