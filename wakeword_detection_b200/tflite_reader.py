"""Dependency-free reader for the `.tflite` flatbuffers the reference ships.

The reference hands its three model files (filter / encode / detect) to
`tflite.Interpreter` (reference: spokestack/models/tensorflow.py:24-31).  That
runtime is not part of this framework; the only thing needed from the files is
the graph structure (to locate each trained tensor robustly) and the tensor
payloads.  This module walks the flatbuffer directly: schema v3, field slots as
listed in SURVEY.md Appendix B.

Nothing here executes a graph.  `oracle/tflite_literal.py` (test infrastructure)
executes graphs op by op on top of this reader; the product path only uses it to
pull trained weights out of a model directory (`weights.py`).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# builtin operator codes that occur in the reference's model files
BUILTIN_NAMES = {
    0: "ADD", 2: "CONCATENATION", 3: "CONV_2D", 9: "FULLY_CONNECTED", 14: "LOGISTIC",
    18: "MUL", 19: "RELU", 22: "RESHAPE", 25: "SOFTMAX", 28: "TANH", 34: "PAD",
    36: "GATHER", 37: "BATCH_TO_SPACE_ND", 38: "SPACE_TO_BATCH_ND", 39: "TRANSPOSE",
    41: "SUB", 45: "STRIDED_SLICE", 49: "SPLIT", 53: "CAST", 55: "MAXIMUM", 58: "LESS",
    65: "SLICE", 70: "EXPAND_DIMS", 73: "LOG", 77: "SHAPE", 82: "REDUCE_MAX", 83: "PACK",
    94: "FILL", 105: "REVERSE_V2", 119: "WHILE", 6: "DEQUANTIZE", 40: "MEAN",
    17: "MAX_POOL_2D", 1: "AVERAGE_POOL_2D", 43: "SQUEEZE", 88: "UNPACK", 61: "LOGICAL_AND",
    74: "SUM", 80: "FAKE_QUANT", 102: "SQUARED_DIFFERENCE", 76: "RSQRT", 42: "DIV",
}

_TENSOR_TYPES = {
    0: np.float32, 1: np.float16, 2: np.int32, 3: np.uint8, 4: np.int64,
    6: np.bool_, 7: np.int16, 9: np.int8, 10: np.float64,
}


class _FB:
    """Minimal flatbuffer table/vector accessor over a bytes object."""

    def __init__(self, buf: bytes) -> None:
        self.b = buf

    def u8(self, o): return self.b[o]
    def i8(self, o): return struct.unpack_from("<b", self.b, o)[0]
    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def i32(self, o): return struct.unpack_from("<i", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def f32(self, o): return struct.unpack_from("<f", self.b, o)[0]

    def indirect(self, o: int) -> int:
        return o + self.u32(o)

    def field(self, tbl: int, slot: int) -> int:
        """Absolute offset of field `slot` in table `tbl`, 0 if absent."""
        vt = tbl - self.i32(tbl)
        vt_len = self.u16(vt)
        pos = 4 + 2 * slot
        if pos >= vt_len:
            return 0
        off = self.u16(vt + pos)
        return tbl + off if off else 0

    def scalar(self, tbl, slot, kind, default=0):
        o = self.field(tbl, slot)
        if not o:
            return default
        return getattr(self, kind)(o)

    def table(self, tbl, slot) -> int:
        o = self.field(tbl, slot)
        return self.indirect(o) if o else 0

    def vector(self, tbl, slot):
        """(start, length) of a vector field; (0, 0) if absent."""
        o = self.field(tbl, slot)
        if not o:
            return 0, 0
        v = self.indirect(o)
        return v + 4, self.u32(v)

    def vec_tables(self, tbl, slot) -> List[int]:
        s, n = self.vector(tbl, slot)
        return [self.indirect(s + 4 * i) for i in range(n)]

    def vec_i32(self, tbl, slot) -> List[int]:
        s, n = self.vector(tbl, slot)
        return list(struct.unpack_from("<%di" % n, self.b, s)) if n else []

    def string(self, tbl, slot) -> str:
        s, n = self.vector(tbl, slot)
        return self.b[s:s + n].decode("utf-8", "replace") if n else ""


@dataclass
class Tensor:
    index: int
    name: str
    shape: List[int]
    dtype: type
    buffer: int
    data: Optional[np.ndarray]  # constant payload or None


@dataclass
class Op:
    code: int
    name: str
    inputs: List[int]
    outputs: List[int]
    options: Dict[str, int] = field(default_factory=dict)


@dataclass
class SubGraph:
    name: str
    tensors: List[Tensor]
    inputs: List[int]
    outputs: List[int]
    ops: List[Op]


@dataclass
class Model:
    version: int
    description: str
    subgraphs: List[SubGraph]

    @property
    def main(self) -> SubGraph:
        return self.subgraphs[0]

    def input_details(self):
        g = self.main
        return [{"index": i, "name": g.tensors[i].name,
                 "shape": np.array(g.tensors[i].shape, dtype=np.int32),
                 "dtype": g.tensors[i].dtype} for i in g.inputs]

    def output_details(self):
        g = self.main
        return [{"index": i, "name": g.tensors[i].name,
                 "shape": np.array(g.tensors[i].shape, dtype=np.int32),
                 "dtype": g.tensors[i].dtype} for i in g.outputs]


def _parse_options(fb: _FB, name: str, opt: int) -> Dict[str, int]:
    if not opt:
        return {}
    if name == "CONV_2D":
        return {"padding": fb.scalar(opt, 0, "i8"), "stride_w": fb.scalar(opt, 1, "i32"),
                "stride_h": fb.scalar(opt, 2, "i32"), "act": fb.scalar(opt, 3, "i8"),
                "dil_w": fb.scalar(opt, 4, "i32", 1), "dil_h": fb.scalar(opt, 5, "i32", 1)}
    if name == "FULLY_CONNECTED":
        return {"act": fb.scalar(opt, 0, "i8"), "keep_num_dims": fb.scalar(opt, 2, "u8")}
    if name in ("ADD", "MUL", "SUB", "DIV"):
        return {"act": fb.scalar(opt, 0, "i8")}
    if name == "WHILE":
        return {"cond": fb.scalar(opt, 0, "i32"), "body": fb.scalar(opt, 1, "i32")}
    if name == "CONCATENATION":
        return {"axis": fb.scalar(opt, 0, "i32"), "act": fb.scalar(opt, 1, "i8")}
    if name == "STRIDED_SLICE":
        return {"begin_mask": fb.scalar(opt, 0, "i32"), "end_mask": fb.scalar(opt, 1, "i32"),
                "ellipsis_mask": fb.scalar(opt, 2, "i32"), "new_axis_mask": fb.scalar(opt, 3, "i32"),
                "shrink_axis_mask": fb.scalar(opt, 4, "i32")}
    if name == "SOFTMAX":
        o = fb.field(opt, 0)
        return {"beta": fb.f32(o) if o else 1.0}
    if name == "SPLIT":
        return {"num_splits": fb.scalar(opt, 0, "i32")}
    if name == "GATHER":
        return {"axis": fb.scalar(opt, 0, "i32")}
    if name == "REDUCE_MAX":
        return {"keep_dims": fb.scalar(opt, 0, "u8")}
    if name == "PACK":
        return {"values_count": fb.scalar(opt, 0, "i32"), "axis": fb.scalar(opt, 1, "i32")}
    if name == "CAST":
        return {}
    return {}


def load(path: str) -> Model:
    """Parse a `.tflite` file into plain Python/numpy structures."""
    with open(path, "rb") as f:
        buf = f.read()
    fb = _FB(buf)
    root = fb.indirect(0)

    # operator codes
    codes = []
    for oc in fb.vec_tables(root, 1):
        dep = fb.scalar(oc, 0, "i8")
        new = fb.scalar(oc, 3, "i32")
        codes.append(max(dep, new))

    # buffers (payload bytes)
    buffers = []
    for b in fb.vec_tables(root, 4):
        s, n = fb.vector(b, 0)
        buffers.append((s, n))

    subgraphs = []
    for sg in fb.vec_tables(root, 2):
        tensors = []
        for ti, t in enumerate(fb.vec_tables(sg, 0)):
            shape = fb.vec_i32(t, 0)
            ttype = fb.scalar(t, 1, "i8")
            bidx = fb.scalar(t, 2, "u32")
            name = fb.string(t, 3)
            dtype = _TENSOR_TYPES.get(ttype, np.float32)
            data = None
            if 0 < bidx < len(buffers):
                s, n = buffers[bidx]
                if n:
                    data = np.frombuffer(buf, dtype=dtype, count=n // np.dtype(dtype).itemsize,
                                         offset=s).copy()
                    if shape:
                        data = data.reshape(shape)
                    elif data.size == 1:
                        data = data.reshape(())
            tensors.append(Tensor(ti, name, shape, dtype, bidx, data))
        ops = []
        for o in fb.vec_tables(sg, 3):
            code = codes[fb.scalar(o, 0, "u32")]
            name = BUILTIN_NAMES.get(code, "OP_%d" % code)
            opt = fb.table(o, 4)
            ops.append(Op(code, name, fb.vec_i32(o, 1), fb.vec_i32(o, 2),
                          _parse_options(fb, name, opt)))
        subgraphs.append(SubGraph(fb.string(sg, 4), tensors, fb.vec_i32(sg, 1),
                                  fb.vec_i32(sg, 2), ops))
    return fold_dequantize(Model(fb.scalar(root, 0, "u32"), fb.string(root, 3), subgraphs))


def fold_dequantize(model: Model) -> Model:
    """fp16-quantised files (`encode-quant.tflite`, reference: wwdetect/CRNN/convert_CRNN_tflite.py:23-37,
    wwdetect/wavenet/wavenet_model.py:149-163) store every trained constant as FLOAT16 behind a DEQUANTIZE op and
    compute in float32.  Folding the op (output tensor := the constant widened to float32) leaves a graph of the same
    shape as the float32 file, so the weight extractors and the literal interpreter read both alike."""
    for g in model.subgraphs:
        keep = []
        for op in g.ops:
            src = g.tensors[op.inputs[0]] if op.name == "DEQUANTIZE" and op.inputs else None
            if src is not None and src.data is not None and src.data.dtype == np.float16:
                dst = g.tensors[op.outputs[0]]
                dst.data = src.data.astype(np.float32)
                dst.dtype = np.float32
                continue
            keep.append(op)
        g.ops = keep
    return model
