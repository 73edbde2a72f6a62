"""ONNX weight import (crazyara_b200/onnx_import.py) against the state-dict import: the same network written
as an ONNX file by tests/onnx_writer.py must give the same blob, byte for byte.  (Parity unpinned: the reference tree holds
no .onnx artefact; see the module header.)"""
import numpy as np
import pytest

from crazyara_b200 import synthetic
from crazyara_b200.onnx_import import import_onnx, read_graph
from crazyara_b200.weights import export_blob
from tests.onnx_writer import write_rise_onnx


@pytest.mark.parametrize("name", ["risev2", "risev33"])
def test_onnx_import_equals_state_dict_import(tmp_path, name):
    arch = synthetic.risev2(34, 81) if name == "risev2" else synthetic.risev33(52, 76)
    sd = synthetic.random_state_dict(arch, 3)
    version = 10 if name == "risev2" else 30
    ref = export_blob(sd, arch, str(tmp_path / "ref.arab"), input_version=version)
    onnx_path = write_rise_onnx(sd, arch, str(tmp_path / "model-v1.0.onnx"))
    nodes, inits = read_graph(onnx_path)
    assert sum(1 for nd in nodes if nd[0] == "Conv") >= 3 * len(arch["kernels"]) + 4 and len(inits) > 50
    got = import_onnx(onnx_path, str(tmp_path / "onnx.arab"))
    assert got["kernels"] == list(arch["kernels"]) and got["c_ops"] == list(arch["c_ops"]) and got["wdl"] == bool(arch["wdl"])
    assert [s or None for s in got["se_types"]] == [s or None for s in arch["se_types"]]
    assert got["in_channels"] == arch["in_channels"] and got["policy_channels"] == arch["policy_channels"]
    a, b = open(ref, "rb").read(), open(tmp_path / "onnx.arab", "rb").read()
    assert a == b


def test_onnx_import_rejects_other_graphs(tmp_path):
    p = tmp_path / "bad.onnx"
    p.write_bytes(b"\x08\x08")
    with pytest.raises(ValueError):
        import_onnx(str(p), str(tmp_path / "x.arab"))
