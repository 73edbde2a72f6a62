"""Weight import from the reference's ONNX artefacts: `<prefix>-v<version>[-bsize-<n>].onnx` as written by
`export_to_onnx` (DeepCrazyhouse/src/training/trainer_agent_pytorch.py:588-650: torch.onnx.export in eval mode, then
onnx-simplifier -- BatchNorm folded into the convolutions) -> ARAB2001 blob for ara_net_create.

No `onnx` package is needed (the image has none): the file is read with a minimal protobuf wire-format reader, only the
fields this import uses (graph.node: op_type, inputs, attributes `group` / `transB`; graph.initializer: dims, data type,
raw / float / int64 data).  The graph is not executed; the network structure is the RISE family's and is recovered from
the ORDER of its Conv / Gemm / MatMul nodes (a torch export lists them in execution order):

    stem conv3x3 | per block: [SE: Gemm, Gemm  or  eca Conv1d] conv1x1, depthwise kxk (group = channels), conv1x1 |
    value head: conv1x1, Gemm(s) | policy head: conv3x3, conv3x3

PARITY UNPINNED: /root/reference holds no .onnx file and the `onnx` / `onnxsim` packages are absent, so this reader is
checked against files written by tests/onnx_writer.py (same node kinds and tensor layouts as the torch exporter emits for
these modules), not against an artefact of the reference itself.
"""
import struct

import numpy as np

from .weights import SE_CODE


# ---------------------------------------------------------------------------------------------- protobuf wire format
def _varint(buf, i):
    r, s = 0, 0
    while True:
        b = buf[i]
        i += 1
        r |= (b & 0x7F) << s
        if not b & 0x80:
            return r, i
        s += 7


def _fields(buf):
    """yields (field number, wire type, value) of one message; length-delimited values as memoryview slices"""
    i, n = 0, len(buf)
    while i < n:
        key, i = _varint(buf, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _varint(buf, i)
        elif w == 1:
            v, i = bytes(buf[i:i + 8]), i + 8
        elif w == 2:
            ln, i = _varint(buf, i)
            v, i = buf[i:i + ln], i + ln
        elif w == 5:
            v, i = bytes(buf[i:i + 4]), i + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {w}")
        yield f, w, v


def _packed_varints(v):
    out, i = [], 0
    while i < len(v):
        x, i = _varint(v, i)
        out.append(x)
    return out


def _tensor(buf):
    """TensorProto -> (name, ndarray float32)"""
    dims, dtype, name, raw, floats, int64s = [], 1, "", None, [], []
    for f, w, v in _fields(buf):
        if f == 1:
            dims += _packed_varints(v) if w == 2 else [v]
        elif f == 2:
            dtype = v
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = bytes(v)
        elif f == 4:
            floats += list(struct.unpack(f"<{len(v) // 4}f", bytes(v))) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f == 7:
            int64s += _packed_varints(v) if w == 2 else [v]
    if dtype == 1:
        a = np.frombuffer(raw, dtype="<f4") if raw is not None else np.asarray(floats, np.float32)
    elif dtype == 10:
        a = np.frombuffer(raw, dtype="<f2").astype(np.float32)
    elif dtype == 11:
        a = np.frombuffer(raw, dtype="<f8").astype(np.float32)
    elif dtype == 7:
        a = (np.frombuffer(raw, dtype="<i8") if raw is not None else np.asarray(int64s, np.int64)).astype(np.float32)
    else:
        return name, None
    return name, np.array(a, dtype=np.float32).reshape(dims if dims else ())


def read_graph(path):
    """-> (nodes [(op_type, inputs, outputs, {attr: int})], initializers {name: ndarray})"""
    model = memoryview(open(path, "rb").read())
    graph = next((v for f, w, v in _fields(model) if f == 7 and w == 2), None)
    if graph is None:
        raise ValueError(f"{path}: no graph in the ONNX model")
    nodes, inits = [], {}
    for f, w, v in _fields(graph):
        if f == 1:
            op, ins, outs, attrs = "", [], [], {}
            for g, gw, gv in _fields(v):
                if g == 1:
                    ins.append(bytes(gv).decode())
                elif g == 2:
                    outs.append(bytes(gv).decode())
                elif g == 4:
                    op = bytes(gv).decode()
                elif g == 5:
                    an, ai, at = "", None, None
                    for h, hw, hv in _fields(gv):
                        if h == 1:
                            an = bytes(hv).decode()
                        elif h == 3:
                            ai = hv
                        elif h == 5 and hw == 2:
                            at = hv  # a tensor attribute (Constant nodes)
                    if ai is not None:
                        attrs[an] = ai
                    if at is not None:
                        attrs[an] = _tensor(at)[1]
            nodes.append((op, ins, outs, attrs))
        elif f == 5:
            name, a = _tensor(v)
            if a is not None:
                inits[name] = a
    for op, ins, outs, attrs in nodes:  # weights kept as Constant nodes instead of initializers
        if op == "Constant" and outs and isinstance(attrs.get("value"), np.ndarray):
            inits[outs[0]] = attrs["value"]
    return nodes, inits


# ---------------------------------------------------------------------------------------------- graph -> blob
def import_onnx(onnx_path, blob_path, input_version=None, channels=256):
    nodes, inits = read_graph(onnx_path)
    order = []  # convolutions and fully-connected layers in execution order
    for op, ins, outs, attrs in nodes:
        if op == "Conv" and len(ins) >= 2 and ins[1] in inits:
            w = inits[ins[1]]
            b = inits[ins[2]] if len(ins) > 2 and ins[2] in inits else np.zeros(w.shape[0], np.float32)
            order.append(("conv", w, b, int(attrs.get("group", 1))))
        elif op in ("Gemm", "MatMul") and len(ins) >= 2 and ins[1] in inits:
            w = inits[ins[1]]
            if op == "MatMul" or not attrs.get("transB", 0):
                w = w.T  # -> [out, in], the layout of torch.nn.Linear.weight
            b = inits[ins[2]] if len(ins) > 2 and ins[2] in inits else np.zeros(w.shape[0], np.float32)
            order.append(("fc", np.ascontiguousarray(w), b, 1))
    if not order or order[0][0] != "conv" or order[0][1].ndim != 4 or order[0][1].shape[2] != 3:
        raise ValueError(f"{onnx_path}: does not start with a 3x3 stem convolution")
    tensors = []
    put = lambda a: tensors.append(np.ascontiguousarray(a, dtype=np.float32).reshape(-1))
    stem_w, stem_b = order[0][1], order[0][2]
    C = stem_w.shape[0]
    if C != channels:
        raise ValueError(f"{onnx_path}: {C} trunk channels, this engine builds {channels}")
    put(stem_w), put(stem_b)
    i, kernels, se_types, c_ops = 1, [], [], []
    while True:
        # squeeze-excitation layers in front of the block (they act on its input)
        se, j = None, i
        fcs = []
        while j < len(order) and (order[j][0] == "fc" or (order[j][0] == "conv" and order[j][1].ndim == 3)):
            fcs.append(order[j])
            j += 1
        # a block = conv1x1 (C -> Cop), depthwise kxk, conv1x1 (Cop -> C)
        if not (j + 2 < len(order) and order[j][0] == "conv" and order[j][1].ndim == 4 and order[j][1].shape[2] == 1 and
                order[j + 1][0] == "conv" and order[j + 1][3] == order[j + 1][1].shape[0] and order[j + 1][3] > 1):
            break
        if fcs:
            if len(fcs) == 2 and fcs[0][0] == "fc":
                se = "ca_se"
                put(fcs[0][1]), put(fcs[1][1])  # (the reference's SE FCs have no bias, builder_util.py:100-116)
            elif len(fcs) == 1 and fcs[0][1].ndim == 3:
                se = "eca_se"
                wc = fcs[0][1]
                put(wc[:, :, wc.shape[2] // 2]), put(fcs[0][2])
            else:
                raise ValueError(f"{onnx_path}: unrecognised squeeze-excitation pattern in front of block {len(kernels)}")
        w1, b1, _ = order[j][1:]
        wd, bd, _ = order[j + 1][1:]
        w2, b2, _ = order[j + 2][1:]
        cop, k = w1.shape[0], wd.shape[2]
        if w1.shape[1] != C or wd.shape[0] != cop or w2.shape[:2] != (C, cop) or k not in (3, 5):
            raise ValueError(f"{onnx_path}: block {len(kernels)} is not a RISE bottleneck block")
        put(w1), put(b1), put(wd), put(bd), put(w2), put(b2)
        kernels.append(int(k)), se_types.append(se), c_ops.append(int(cop))
        i = j + 3
    if not kernels:
        raise ValueError(f"{onnx_path}: no bottleneck block found")
    # value head: conv1x1 (C -> 8) + FC layers; policy head: conv3x3 (C -> C), conv3x3 (C -> P)
    rest = order[i:]
    vconv = [r for r in rest if r[0] == "conv" and r[1].ndim == 4 and r[1].shape[2] == 1]
    heads = [r for r in rest if r[0] == "conv" and r[1].ndim == 4 and r[1].shape[2] == 3]
    fcs = [r for r in rest if r[0] == "fc"]
    if len(vconv) != 1 or vconv[0][1].shape[1] != C:
        raise ValueError(f"{onnx_path}: value head (one 1x1 convolution) not found behind the tower")
    if len(heads) != 2 or heads[0][1].shape[:2] != (C, C) or heads[1][1].shape[1] != C:
        raise ValueError(f"{onnx_path}: policy head (two 3x3 convolutions) not found behind the tower")
    put(vconv[0][1]), put(vconv[0][2])
    # value head variants (builder_util.py:246-330): FC 512 -> 256 -> 1 (tanh), or the WDL head FC -> 3 with the plies FC -> 1
    wdl = len(fcs) == 2 and fcs[0][1].shape[0] == 3
    if wdl:
        put(fcs[0][1]), put(fcs[0][2]), put(fcs[1][1]), put(fcs[1][2])
    elif len(fcs) == 2:
        put(fcs[0][1]), put(fcs[0][2]), put(fcs[1][1]), put(fcs[1][2])
    else:
        raise ValueError(f"{onnx_path}: {len(fcs)} fully-connected layers in the value head (expected 2)")
    put(heads[0][1]), put(heads[0][2])
    put(heads[1][1])
    arch = dict(name=f"rise_{len(kernels)}b", in_channels=int(stem_w.shape[1]), policy_channels=int(heads[1][1].shape[0]),
                channels=int(C), kernels=kernels, se_types=se_types, c_ops=c_ops, wdl=bool(wdl), value_channels=int(vconv[0][1].shape[0]),
                value_fc=256)
    if input_version is None:
        input_version = {34: 10, 63: 10, 39: 10, 51: 20, 52: 30, 64: 30, 80: 30}.get(arch["in_channels"], 10)
    with open(blob_path, "wb") as f:
        f.write(b"ARAB2001")
        f.write(struct.pack("<8i", arch["in_channels"], arch["policy_channels"], len(kernels), 256, 8, 256, 1 if wdl else 0, input_version))
        for k, se, cop in zip(kernels, se_types, c_ops):
            f.write(struct.pack("<3i", cop, k, SE_CODE[se]))
        for t in tensors:
            f.write(struct.pack("<q", t.size))
            f.write(t.tobytes())
    return arch


if __name__ == "__main__":
    import sys
    if len(sys.argv) < 3:
        raise SystemExit("usage: python -m crazyara_b200.onnx_import <model.onnx> <out.arab> [input_version]")
    a = import_onnx(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    print(f"{sys.argv[2]}: {len(a['kernels'])} blocks, {a['in_channels']} -> {a['policy_channels']}x64, wdl={a['wdl']}")
