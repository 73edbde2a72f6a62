"""Weight import: a reference-format checkpoint converts without naming the architecture."""
import numpy as np
import torch

from crazyara_b200.weights import arch_from_state_dict, export_blob, import_checkpoint
from oracle import net as onet


def test_architecture_is_read_off_the_state_dict():
    for arch in (onet.arch_risev2(34, 81), onet.arch_risev33(52, 76, True), onet.arch_risev2(63, 84)):
        sd = onet.make_state_dict(arch, 1)
        got = arch_from_state_dict(sd)
        for k in ("in_channels", "policy_channels", "channels", "kernels", "se_types", "c_ops", "wdl"):
            assert got[k] == arch[k], k


def test_checkpoint_import_equals_direct_export(tmp_path):
    arch = onet.arch_risev33(52, 76, True)
    sd = onet.make_state_dict(arch, 2)
    ck = tmp_path / "model.tar"
    torch.save({"model_state_dict": {"module." + k: v for k, v in sd.items()}, "epoch": 3}, ck)
    a = import_checkpoint(str(ck), str(tmp_path / "a.arab"))
    export_blob(sd, arch, str(tmp_path / "b.arab"), input_version=30)
    assert a["kernels"] == arch["kernels"]
    assert (tmp_path / "a.arab").read_bytes() == (tmp_path / "b.arab").read_bytes()
