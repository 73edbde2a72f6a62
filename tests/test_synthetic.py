"""The product's random-network generator (crazyara_b200/synthetic.py) against the oracle's architecture tables, and the
rule that the product never reaches into oracle/."""
import ast
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_architecture_tables_and_parameter_shapes_match_the_oracle(tmp_path):
    from crazyara_b200 import synthetic
    from crazyara_b200.weights import arch_from_state_dict, export_blob
    from oracle import net as onet
    for mine, ref in ((synthetic.risev2(34, 81), onet.arch_risev2(34, 81)),
                      (synthetic.risev2(63, 84), onet.arch_risev2(63, 84)),
                      (synthetic.risev33(52, 76), onet.arch_risev33(52, 76)),
                      (synthetic.risev33(64, 81, wdl=False), onet.arch_risev33(64, 81, wdl=False))):
        assert mine == ref
        sd, sd_ref = synthetic.random_state_dict(mine, 3), onet.make_state_dict(ref, 3)
        assert sd.keys() == sd_ref.keys()
        assert all(sd[k].shape == sd_ref[k].shape and sd[k].dtype == np.float32 for k in sd)
        # the generated parameters describe the architecture they were made for, and the network stays in fp16 range
        got = arch_from_state_dict(sd)
        assert all(got[k] == mine[k] for k in ("kernels", "se_types", "c_ops", "wdl", "in_channels", "policy_channels"))
        x = np.random.default_rng(0).random((2, mine["in_channels"], 8, 8), dtype=np.float32)
        out = onet.forward(sd, mine, x)
        assert np.isfinite(out["prob"]).all() and abs(float(out["prob"].sum()) - 2.0) < 1e-3
        assert os.path.getsize(export_blob(sd, mine, str(tmp_path / (mine["name"] + ".arab")), input_version=10)) > 1 << 20
    assert not np.array_equal(synthetic.random_state_dict(mine, 1)["policy_head.body.3.weight"],
                              synthetic.random_state_dict(mine, 2)["policy_head.body.3.weight"])


def _imports(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield node.lineno, a.name
        elif isinstance(node, ast.ImportFrom) and node.module:
            yield node.lineno, node.module


def test_product_code_never_imports_the_oracle():
    """oracle/ is the checker: only tests/, __graft_entry__.smoke() and bench.py's CPU arm may use it."""
    offenders = []
    for base, _, files in os.walk(os.path.join(ROOT, "crazyara_b200")):
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(base, f)
                offenders += [(p, ln) for ln, mod in _imports(p) if mod.split(".")[0] == "oracle"]
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            p = os.path.join(ROOT, "tools", f)
            offenders += [(p, ln) for ln, mod in _imports(p) if mod.split(".")[0] == "oracle"]
    assert offenders == []
    # bench.py: the oracle appears inside cpu_arm() only
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    for node in tree.body:
        inside = isinstance(node, ast.FunctionDef) and node.name == "cpu_arm"
        for sub in ast.walk(node):
            mods = []
            if isinstance(sub, ast.Import):
                mods = [a.name for a in sub.names]
            elif isinstance(sub, ast.ImportFrom) and sub.module:
                mods = [sub.module]
            assert inside or all(m.split(".")[0] != "oracle" for m in mods), (getattr(node, "name", "?"), mods)
    # and no C/CUDA source of the library includes anything from oracle/ (comments may cite it)
    for base in (os.path.join(ROOT, "crazyara_b200", "csrc"), os.path.join(ROOT, "crazyara_b200", "host"), os.path.join(ROOT, "include")):
        for f in os.listdir(base):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                for line in open(os.path.join(base, f)):
                    assert not (line.lstrip().startswith("#include") and "oracle" in line), (f, line)
