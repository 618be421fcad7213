"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

Independent `.onnx` decoder for the oracles.  The product decodes the protobuf wire format by hand
(`metagenomic-deepfri_b200/onnx_lite.py`, and `csrc/onnx_load.cpp` behind the C ABI); so that a decoding
mistake cannot hide on both sides of a parity test, the oracles parse the same bytes with Google's own
protobuf runtime (`google.protobuf`, present in the image) over message descriptors declared here from the
public `onnx.proto` (IR version 8; only the messages / fields an inference graph uses).  Nothing in this file
shares code with the product.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, List

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_OPT, _REP = _F.LABEL_OPTIONAL, _F.LABEL_REPEATED


def _build_pool():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "mdf_oracle_onnx_subset.proto"
    fd.package = "mdf_oracle_onnx"
    fd.syntax = "proto2"

    def msg(name, fields, parent=None):
        m = (parent.nested_type if parent is not None else fd.message_type).add()
        m.name = name
        for fname, num, ftype, label, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, ftype, label
            if tname:
                f.type_name = ".mdf_oracle_onnx." + tname
        return m

    S, I64, I32, FL, DB, BY, MSG, U64 = (_F.TYPE_STRING, _F.TYPE_INT64, _F.TYPE_INT32, _F.TYPE_FLOAT, _F.TYPE_DOUBLE,
                                         _F.TYPE_BYTES, _F.TYPE_MESSAGE, _F.TYPE_UINT64)
    msg("TensorProto", [
        ("dims", 1, I64, _REP, None), ("data_type", 2, I32, _OPT, None), ("float_data", 4, FL, _REP, None),
        ("int32_data", 5, I32, _REP, None), ("string_data", 6, BY, _REP, None), ("int64_data", 7, I64, _REP, None),
        ("name", 8, S, _OPT, None), ("raw_data", 9, BY, _OPT, None), ("double_data", 10, DB, _REP, None),
        ("uint64_data", 11, U64, _REP, None), ("doc_string", 12, S, _OPT, None), ("data_location", 14, I32, _OPT, None)])
    shp = msg("TensorShapeProto", [("dim", 1, MSG, _REP, "TensorShapeProto.Dimension")])
    msg("Dimension", [("dim_value", 1, I64, _OPT, None), ("dim_param", 2, S, _OPT, None), ("denotation", 3, S, _OPT, None)], shp)
    typ = msg("TypeProto", [("tensor_type", 1, MSG, _OPT, "TypeProto.Tensor"), ("denotation", 6, S, _OPT, None)])
    msg("Tensor", [("elem_type", 1, I32, _OPT, None), ("shape", 2, MSG, _OPT, "TensorShapeProto")], typ)
    msg("ValueInfoProto", [("name", 1, S, _OPT, None), ("type", 2, MSG, _OPT, "TypeProto"), ("doc_string", 3, S, _OPT, None)])
    msg("AttributeProto", [
        ("name", 1, S, _OPT, None), ("f", 2, FL, _OPT, None), ("i", 3, I64, _OPT, None), ("s", 4, BY, _OPT, None),
        ("t", 5, MSG, _OPT, "TensorProto"), ("g", 6, MSG, _OPT, "GraphProto"), ("floats", 7, FL, _REP, None),
        ("ints", 8, I64, _REP, None), ("strings", 9, BY, _REP, None), ("tensors", 10, MSG, _REP, "TensorProto"),
        ("doc_string", 13, S, _OPT, None), ("type", 20, I32, _OPT, None), ("ref_attr_name", 21, S, _OPT, None)])
    msg("NodeProto", [
        ("input", 1, S, _REP, None), ("output", 2, S, _REP, None), ("name", 3, S, _OPT, None), ("op_type", 4, S, _OPT, None),
        ("attribute", 5, MSG, _REP, "AttributeProto"), ("doc_string", 6, S, _OPT, None), ("domain", 7, S, _OPT, None)])
    msg("GraphProto", [
        ("node", 1, MSG, _REP, "NodeProto"), ("name", 2, S, _OPT, None), ("initializer", 5, MSG, _REP, "TensorProto"),
        ("doc_string", 10, S, _OPT, None), ("input", 11, MSG, _REP, "ValueInfoProto"), ("output", 12, MSG, _REP, "ValueInfoProto"),
        ("value_info", 13, MSG, _REP, "ValueInfoProto")])
    msg("OperatorSetIdProto", [("domain", 1, S, _OPT, None), ("version", 2, I64, _OPT, None)])
    msg("ModelProto", [
        ("ir_version", 1, I64, _OPT, None), ("producer_name", 2, S, _OPT, None), ("producer_version", 3, S, _OPT, None),
        ("domain", 4, S, _OPT, None), ("model_version", 5, I64, _OPT, None), ("doc_string", 6, S, _OPT, None),
        ("graph", 7, MSG, _OPT, "GraphProto"), ("opset_import", 8, MSG, _REP, "OperatorSetIdProto")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


_POOL = _build_pool()
ModelProto = message_factory.GetMessageClass(_POOL.FindMessageTypeByName("mdf_oracle_onnx.ModelProto"))

# TensorProto.DataType -> NumPy
NP_OF = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16,
         11: np.float64, 12: np.uint32, 13: np.uint64}


def tensor_to_numpy(t) -> np.ndarray:
    if t.data_location == 1:
        raise ValueError(f"tensor {t.name!r} uses external data")
    if t.data_type not in NP_OF:
        raise ValueError(f"tensor {t.name!r}: unsupported ONNX data_type {t.data_type}")
    dt = np.dtype(NP_OF[t.data_type])
    dims = [int(d) for d in t.dims]
    if t.HasField("raw_data"):
        arr = np.frombuffer(t.raw_data, dtype=dt.newbyteorder("<")).astype(dt)
    elif t.data_type == 1:
        arr = np.asarray(list(t.float_data), dt)
    elif t.data_type == 11:
        arr = np.asarray(list(t.double_data), dt)
    elif t.data_type == 7:
        arr = np.asarray(list(t.int64_data), dt)
    elif t.data_type in (12, 13):
        arr = np.asarray(list(t.uint64_data), dt)
    elif t.data_type == 10:
        arr = np.asarray(list(t.int32_data), np.uint16).view(np.float16)      # fp16 bit patterns travel in int32_data
    else:
        arr = np.asarray(list(t.int32_data), dt)
    return arr.reshape(dims)


def _attr_value(a) -> Any:
    # AttributeProto.AttributeType: FLOAT 1, INT 2, STRING 3, TENSOR 4, GRAPH 5, FLOATS 6, INTS 7, STRINGS 8
    t = a.type
    if t == 1 or (t == 0 and a.HasField("f")):
        return float(a.f)
    if t == 2 or (t == 0 and a.HasField("i")):
        return int(a.i)
    if t == 3 or (t == 0 and a.HasField("s")):
        return a.s.decode("utf-8", "replace")
    if t == 4 or (t == 0 and a.HasField("t")):
        return tensor_to_numpy(a.t)
    if t == 6:
        return [float(x) for x in a.floats]
    if t == 7:
        return [int(x) for x in a.ints]
    if t == 8:
        return [x.decode("utf-8", "replace") for x in a.strings]
    if len(a.ints):
        return [int(x) for x in a.ints]
    if len(a.floats):
        return [float(x) for x in a.floats]
    return None


def load(path: str) -> SimpleNamespace:
    """-> namespace(graph=namespace(nodes, initializers, inputs, outputs), opset, ir_version); nodes carry
    op_type / inputs / outputs / name / attrs, value infos carry name / elem_type / shape."""
    with open(path, "rb") as fh:
        data = fh.read()
    m = ModelProto()
    m.ParseFromString(data)
    if not m.HasField("graph"):
        raise ValueError(f"{path}: not an ONNX ModelProto (no graph)")
    g = m.graph
    init: Dict[str, np.ndarray] = {t.name: tensor_to_numpy(t) for t in g.initializer}
    nodes: List[SimpleNamespace] = []
    for n in g.node:
        nodes.append(SimpleNamespace(op_type=n.op_type, inputs=list(n.input), outputs=list(n.output), name=n.name,
                                     attrs={a.name: _attr_value(a) for a in n.attribute}))

    def vinfo(v):
        shape = []
        for d in v.type.tensor_type.shape.dim:
            shape.append(int(d.dim_value) if d.HasField("dim_value") else (d.dim_param if d.HasField("dim_param") else None))
        return SimpleNamespace(name=v.name, elem_type=int(v.type.tensor_type.elem_type), shape=tuple(shape))
    opset = 0
    for o in m.opset_import:
        if o.domain in ("", "ai.onnx"):
            opset = int(o.version)
    graph = SimpleNamespace(name=g.name, nodes=nodes, initializers=init,
                            inputs=[vinfo(v) for v in g.input if v.name not in init],
                            outputs=[vinfo(v) for v in g.output])
    return SimpleNamespace(graph=graph, opset=opset, ir_version=int(m.ir_version), producer_name=m.producer_name)
