"""Minimal ONNX protobuf reader/writer (no `onnx`, no `protoc`).

The reference hands `<model>.onnx` files to onnxruntime (`mDeepFRI/predict.pyx:62-73`);
neither `onnx` nor `onnxruntime` exist in this image, so the wire format is decoded by
hand.  Only the subset of `onnx.proto` needed for tf2onnx-style inference graphs is
covered: ModelProto / GraphProto / NodeProto / AttributeProto / TensorProto /
ValueInfoProto.  Field numbers follow the public onnx.proto (IR version 8).

This module only moves bytes <-> Python objects.  It never interprets a graph.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

# TensorProto.DataType
FLOAT, UINT8, INT8, INT32, INT64, BOOL, FLOAT16, DOUBLE = 1, 2, 3, 6, 7, 9, 10, 11
_NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, INT32: np.int32,
          INT64: np.int64, BOOL: np.bool_, FLOAT16: np.float16, DOUBLE: np.float64}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}

# AttributeProto.AttributeType
A_FLOAT, A_INT, A_STRING, A_TENSOR, A_FLOATS, A_INTS, A_STRINGS = 1, 2, 3, 4, 6, 7, 8


@dataclass
class Node:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    name: str = ""
    attrs: Dict[str, Any] = field(default_factory=dict)


@dataclass
class ValueInfo:
    name: str
    elem_type: int = FLOAT
    shape: Tuple[Any, ...] = ()      # ints or str (symbolic) or None


@dataclass
class Graph:
    name: str = "graph"
    nodes: List[Node] = field(default_factory=list)
    initializers: Dict[str, np.ndarray] = field(default_factory=dict)
    inputs: List[ValueInfo] = field(default_factory=list)
    outputs: List[ValueInfo] = field(default_factory=list)


@dataclass
class Model:
    graph: Graph
    ir_version: int = 8
    opset: int = 15
    producer_name: str = ""
    producer_version: str = ""


# ----------------------------------------------------------------------------- wire decode
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) for one message body."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            if len(v) != ln:
                raise ValueError("truncated length-delimited field")
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(wt: int, v) -> List[int]:
    if wt == 0:
        return [_signed64(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed64(x))
    return out


def _parse_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = FLOAT
    name = ""
    raw: Optional[bytes] = None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(wt, v)
        elif fno == 2:
            dtype = v
        elif fno == 4:   # float_data
            if wt == 5:
                floats.append(struct.unpack("<f", v)[0])
            else:
                floats += list(np.frombuffer(v, dtype="<f4"))
        elif fno in (5, 7):  # int32_data / int64_data
            ints += _packed_varints(wt, v)
        elif fno == 8:
            name = v.decode("utf-8")
        elif fno == 9:
            raw = bytes(v)
        elif fno == 10:  # double_data
            floats += list(np.frombuffer(v, dtype="<f8")) if wt == 2 else [struct.unpack("<d", v)[0]]
    if dtype not in _NP_OF:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data_type {dtype}")
    npdt = np.dtype(_NP_OF[dtype])
    if raw is not None:
        arr = np.frombuffer(raw, dtype=npdt.newbyteorder("<")).astype(npdt)
    elif dtype in (FLOAT, DOUBLE):
        arr = np.asarray(floats, dtype=npdt)
    else:
        arr = np.asarray(ints, dtype=npdt)
    return name, arr.reshape(dims) if dims else arr.reshape(())


def _parse_attr(buf: bytes) -> Tuple[str, Any]:
    name, atype = "", 0
    f = i = s = t = None
    floats: List[float] = []
    ints: List[int] = []
    strings: List[bytes] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode("utf-8")
        elif fno == 2:
            f = struct.unpack("<f", v)[0]
        elif fno == 3:
            i = _signed64(v)
        elif fno == 4:
            s = bytes(v)
        elif fno == 5:
            t = _parse_tensor(v)[1]
        elif fno == 7:
            floats += list(np.frombuffer(v, dtype="<f4")) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 8:
            ints += _packed_varints(wt, v)
        elif fno == 9:
            strings.append(bytes(v))
        elif fno == 20:
            atype = v
    if atype == A_FLOAT or (atype == 0 and f is not None):
        return name, float(f)
    if atype == A_INT or (atype == 0 and i is not None):
        return name, int(i)
    if atype == A_STRING or (atype == 0 and s is not None):
        return name, s.decode("utf-8", "replace")
    if atype == A_TENSOR or (atype == 0 and t is not None):
        return name, t
    if atype == A_FLOATS:
        return name, [float(x) for x in floats]
    if atype == A_INTS:
        return name, [int(x) for x in ints]
    if atype == A_STRINGS:
        return name, [x.decode("utf-8", "replace") for x in strings]
    return name, ints or floats or None


def _parse_node(buf: bytes) -> Node:
    n = Node("", [], [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n.inputs.append(v.decode("utf-8"))
        elif fno == 2:
            n.outputs.append(v.decode("utf-8"))
        elif fno == 3:
            n.name = v.decode("utf-8")
        elif fno == 4:
            n.op_type = v.decode("utf-8")
        elif fno == 5:
            k, val = _parse_attr(v)
            n.attrs[k] = val
    return n


def _parse_value_info(buf: bytes) -> ValueInfo:
    vi = ValueInfo("")
    for fno, wt, v in _fields(buf):
        if fno == 1:
            vi.name = v.decode("utf-8")
        elif fno == 2:  # TypeProto
            for f2, _, v2 in _fields(v):
                if f2 != 1:      # tensor_type
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        vi.elem_type = v3
                    elif f3 == 2:  # TensorShapeProto
                        dims: List[Any] = []
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            d: Any = None
                            for f5, _, v5 in _fields(v4):
                                if f5 == 1:
                                    d = _signed64(v5)
                                elif f5 == 2:
                                    d = v5.decode("utf-8")
                            dims.append(d)
                        vi.shape = tuple(dims)
    return vi


def _parse_graph(buf: bytes) -> Graph:
    g = Graph()
    for fno, wt, v in _fields(buf):
        if fno == 1:
            g.nodes.append(_parse_node(v))
        elif fno == 2:
            g.name = v.decode("utf-8")
        elif fno == 5:
            name, arr = _parse_tensor(v)
            g.initializers[name] = arr
        elif fno == 11:
            g.inputs.append(_parse_value_info(v))
        elif fno == 12:
            g.outputs.append(_parse_value_info(v))
    # graph inputs that are really initializers (IR < 4 style) are not runtime inputs
    g.inputs = [vi for vi in g.inputs if vi.name not in g.initializers]
    return g


def loads(buf: bytes) -> Model:
    graph = None
    m = Model(Graph())
    for fno, wt, v in _fields(buf):
        if fno == 1:
            m.ir_version = v
        elif fno == 2:
            m.producer_name = v.decode("utf-8")
        elif fno == 3:
            m.producer_version = v.decode("utf-8")
        elif fno == 7:
            graph = _parse_graph(v)
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    dom = v2.decode("utf-8")
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                m.opset = ver
    if graph is None:
        raise ValueError("not an ONNX ModelProto: no graph field")
    m.graph = graph
    return m


def load(path: str) -> Model:
    with open(path, "rb") as fh:     # FileNotFoundError propagates (predict.pyi:70-72)
        data = fh.read()
    try:
        return loads(data)
    except (IndexError, ValueError, struct.error) as e:
        raise RuntimeError(f"{path}: cannot parse as ONNX protobuf: {e}") from e


# ----------------------------------------------------------------------------- wire encode
def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(fno: int, wt: int) -> bytes:
    return _enc_varint((fno << 3) | wt)


def _ld(fno: int, payload: bytes) -> bytes:
    return _key(fno, 2) + _enc_varint(len(payload)) + payload


def _vi(fno: int, v: int) -> bytes:
    return _key(fno, 0) + _enc_varint(v)


def _enc_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)
    if arr.dtype not in _DT_OF:
        raise ValueError(f"unsupported dtype {arr.dtype}")
    out = b"".join(_vi(1, int(d)) for d in arr.shape)
    out += _vi(2, _DT_OF[arr.dtype])
    out += _ld(8, name.encode())
    out += _ld(9, np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes())
    return out


def _enc_attr(name: str, val: Any) -> bytes:
    out = _ld(1, name.encode())
    if isinstance(val, bool):
        val = int(val)
    if isinstance(val, float):
        out += _key(2, 5) + struct.pack("<f", val) + _vi(20, A_FLOAT)
    elif isinstance(val, int):
        out += _vi(3, val) + _vi(20, A_INT)
    elif isinstance(val, str):
        out += _ld(4, val.encode()) + _vi(20, A_STRING)
    elif isinstance(val, np.ndarray):
        out += _ld(5, _enc_tensor("", val)) + _vi(20, A_TENSOR)
    elif isinstance(val, (list, tuple)) and val and isinstance(val[0], float):
        out += _ld(7, np.asarray(val, "<f4").tobytes()) + _vi(20, A_FLOATS)
    elif isinstance(val, (list, tuple)) and val and isinstance(val[0], str):
        out += b"".join(_ld(9, s.encode()) for s in val) + _vi(20, A_STRINGS)
    elif isinstance(val, (list, tuple)):
        out += _ld(8, b"".join(_enc_varint(int(x)) for x in val)) + _vi(20, A_INTS)
    else:
        raise ValueError(f"unsupported attribute {name}={val!r}")
    return out


def _enc_node(n: Node) -> bytes:
    out = b"".join(_ld(1, s.encode()) for s in n.inputs)
    out += b"".join(_ld(2, s.encode()) for s in n.outputs)
    out += _ld(3, n.name.encode()) + _ld(4, n.op_type.encode())
    out += b"".join(_ld(5, _enc_attr(k, v)) for k, v in n.attrs.items())
    return out


def _enc_value_info(vi: ValueInfo) -> bytes:
    dims = b""
    for d in vi.shape:
        if isinstance(d, str):
            dims += _ld(1, _ld(2, d.encode()))
        elif d is None:
            dims += _ld(1, b"")
        else:
            dims += _ld(1, _vi(1, int(d)))
    tensor_type = _vi(1, vi.elem_type) + _ld(2, dims)
    return _ld(1, vi.name.encode()) + _ld(2, _ld(1, tensor_type))


def dumps(m: Model) -> bytes:
    g = m.graph
    gb = b"".join(_ld(1, _enc_node(n)) for n in g.nodes)
    gb += _ld(2, g.name.encode())
    gb += b"".join(_ld(5, _enc_tensor(k, v)) for k, v in g.initializers.items())
    gb += b"".join(_ld(11, _enc_value_info(v)) for v in g.inputs)
    gb += b"".join(_ld(12, _enc_value_info(v)) for v in g.outputs)
    out = _vi(1, m.ir_version)
    out += _ld(2, m.producer_name.encode()) + _ld(3, m.producer_version.encode())
    out += _ld(7, gb)
    out += _ld(8, _ld(1, b"") + _vi(2, m.opset))
    return out


def save(m: Model, path: str) -> None:
    with open(path, "wb") as fh:
        fh.write(dumps(m))
