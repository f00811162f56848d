"""ORACLE (test infrastructure, not product): literal op-by-op execution of the
reference's `.tflite` graphs in numpy float32.

PARITY STATUS: *unpinned by the reference*.  The reference ships no tests, golden
vectors or fixtures and its arithmetic lives in the
third-party `tensorflow==2.4.0` TFLite interpreter (reference requirements.txt:3;
call sites spokestack/models/tensorflow.py:43-51), which cannot be imported in
this environment.  This file therefore restates the *published semantics of the
TFLite float builtin ops* and executes the shipped graphs (weights + op list are
the only specification) exactly in file order, including the WHILE control flow
of the CRNN encoder.  `oracle/restated.py` (the closed-form model the CUDA kernels
mirror) is validated against this file; both are pinned against the known-answer
values recorded in SURVEY.md Appendix A4.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wakeword_detection_b200 import tflite_reader as tr  # noqa: E402

F32 = np.float32


def _act(x, code):
    if code == 0:
        return x
    if code == 1:  # RELU
        return np.maximum(x, F32(0))
    if code == 3:  # RELU6
        return np.minimum(np.maximum(x, F32(0)), F32(6))
    raise NotImplementedError("fused activation %d" % code)


def _conv2d(x, w, b, opt):
    """TFLite CONV_2D float: NHWC input, OHWI weights, SAME/VALID padding."""
    n, ih, iw, ic = x.shape
    oc, kh, kw, _ = w.shape
    sh, sw = opt["stride_h"], opt["stride_w"]
    dh, dw = opt.get("dil_h", 1) or 1, opt.get("dil_w", 1) or 1
    ekh, ekw = (kh - 1) * dh + 1, (kw - 1) * dw + 1
    if opt["padding"] == 0:  # SAME
        oh, ow = -(-ih // sh), -(-iw // sw)
        ph = max((oh - 1) * sh + ekh - ih, 0)
        pw = max((ow - 1) * sw + ekw - iw, 0)
        pt, pl = ph // 2, pw // 2
        x = np.pad(x, ((0, 0), (pt, ph - pt), (pl, pw - pl), (0, 0)))
    else:
        oh, ow = (ih - ekh) // sh + 1, (iw - ekw) // sw + 1
    out = np.zeros((n, oh, ow, oc), dtype=F32)
    wmat = w.reshape(oc, kh * kw * ic).astype(F32)
    for y in range(oh):
        for xx in range(ow):
            patch = x[:, y * sh:y * sh + ekh:dh, xx * sw:xx * sw + ekw:dw, :]
            out[:, y, xx, :] = patch.reshape(n, -1).astype(F32) @ wmat.T
    if b is not None:
        out = out + b.astype(F32)
    return _act(out.astype(F32), opt["act"])


def _strided_slice(x, begin, end, strides, opt):
    idx = []
    for d in range(len(begin)):
        b, e, s = int(begin[d]), int(end[d]), int(strides[d])
        dim = x.shape[d]
        if opt["shrink_axis_mask"] >> d & 1:
            idx.append(b if b >= 0 else b + dim)
            continue
        bb = None if opt["begin_mask"] >> d & 1 else b
        ee = None if opt["end_mask"] >> d & 1 else e
        idx.append(slice(bb, ee, s))
    return x[tuple(idx)]


def _space_to_batch(x, block, pads):
    # x: [N, spatial..., C]
    nb = len(block)
    padcfg = [(0, 0)] + [(int(pads[i][0]), int(pads[i][1])) for i in range(nb)] + \
             [(0, 0)] * (x.ndim - 1 - nb)
    x = np.pad(x, padcfg)
    n = x.shape[0]
    sp = x.shape[1:1 + nb]
    rest = x.shape[1 + nb:]
    shp = [n]
    for i in range(nb):
        shp += [sp[i] // int(block[i]), int(block[i])]
    x = x.reshape(shp + list(rest))
    # move block dims to the front: [b0, b1.., N, s0/b0, s1/b1.., rest]
    perm = [2 + 2 * i for i in range(nb)] + [0] + [1 + 2 * i for i in range(nb)] + \
           list(range(1 + 2 * nb, x.ndim))
    x = x.transpose(perm)
    outshape = [n * int(np.prod(block))] + [sp[i] // int(block[i]) for i in range(nb)] + list(rest)
    return x.reshape(outshape)


def _batch_to_space(x, block, crops):
    nb = len(block)
    prod = int(np.prod(block))
    n = x.shape[0] // prod
    sp = x.shape[1:1 + nb]
    rest = x.shape[1 + nb:]
    x = x.reshape([int(b) for b in block] + [n] + list(sp) + list(rest))
    # -> [N, s0, b0, s1, b1, ..., rest]
    perm = [nb]
    for i in range(nb):
        perm += [nb + 1 + i, i]
    perm += list(range(2 * nb + 1, x.ndim))
    x = x.transpose(perm)
    x = x.reshape([n] + [sp[i] * int(block[i]) for i in range(nb)] + list(rest))
    idx = [slice(None)]
    for i in range(nb):
        c0, c1 = int(crops[i][0]), int(crops[i][1])
        idx.append(slice(c0, x.shape[1 + i] - c1))
    return x[tuple(idx)]


class LiteralInterpreter:
    """Executes a parsed model literally.  Mirrors the call protocol of the
    reference's `TFLiteModel.__call__` (spokestack/models/tensorflow.py:33-51):
    positional input arrays in, list of output arrays out."""

    def __init__(self, path: str) -> None:
        self.model = tr.load(path)
        self.input_details = self.model.input_details()
        self.output_details = self.model.output_details()

    def __call__(self, *args) -> List[np.ndarray]:
        return self._run(0, [np.asarray(a) for a in args])

    # ------------------------------------------------------------------
    def _run(self, sg_index: int, inputs: List[np.ndarray]) -> List[np.ndarray]:
        g = self.model.subgraphs[sg_index]
        val: Dict[int, np.ndarray] = {}
        for t in g.tensors:
            if t.data is not None:
                val[t.index] = t.data
        for i, a in zip(g.inputs, inputs):
            val[i] = a
        for op in g.ops:
            ins = [val[i] if i >= 0 else None for i in op.inputs]
            outs = self._exec(g, op, ins)
            for i, o in zip(op.outputs, outs):
                val[i] = o
        return [val[i] for i in g.outputs]

    def _exec(self, g, op, x):
        n, o = op.name, op.options
        if n == "CONV_2D":
            return [_conv2d(x[0].astype(F32), x[1], x[2] if len(x) > 2 else None, o)]
        if n == "FULLY_CONNECTED":
            y = x[0].reshape(-1, x[1].shape[1]).astype(F32) @ x[1].T.astype(F32)
            if len(x) > 2 and x[2] is not None:
                y = y + x[2]
            return [_act(y.astype(F32), o["act"])]
        if n == "ADD":
            return [_act(np.add(x[0], x[1]), o.get("act", 0))]
        if n == "SUB":
            return [_act(np.subtract(x[0], x[1]), o.get("act", 0))]
        if n == "MUL":
            return [_act(np.multiply(x[0], x[1]), o.get("act", 0))]
        if n == "MAXIMUM":
            return [np.maximum(x[0], x[1])]
        if n == "LOG":
            return [np.log(x[0].astype(F32)).astype(F32)]
        if n == "LOGISTIC":
            return [(F32(1) / (F32(1) + np.exp(-x[0].astype(F32)))).astype(F32)]
        if n == "TANH":
            return [np.tanh(x[0].astype(F32)).astype(F32)]
        if n == "RELU":
            return [np.maximum(x[0], F32(0))]
        if n == "SOFTMAX":
            z = x[0].astype(F32) * F32(o.get("beta", 1.0))
            z = z - z.max(axis=-1, keepdims=True)
            e = np.exp(z).astype(F32)
            return [(e / e.sum(axis=-1, keepdims=True)).astype(F32)]
        if n == "RESHAPE":
            shape = [int(v) for v in np.asarray(x[1]).reshape(-1)]
            return [x[0].reshape(shape)]
        if n == "TRANSPOSE":
            return [np.transpose(x[0], [int(v) for v in x[1]])]
        if n == "SHAPE":
            return [np.array(x[0].shape, dtype=np.int32)]
        if n == "STRIDED_SLICE":
            return [np.asarray(_strided_slice(x[0], x[1], x[2], x[3], o))]
        if n == "PACK":
            return [np.stack([np.asarray(v) for v in x], axis=o.get("axis", 0))]
        if n == "FILL":
            dims = [int(v) for v in np.asarray(x[0]).reshape(-1)]
            return [np.full(dims, x[1], dtype=np.asarray(x[1]).dtype)]
        if n == "REVERSE_V2":
            return [np.flip(x[0], axis=tuple(int(v) for v in np.asarray(x[1]).reshape(-1)))]
        if n == "CONCATENATION":
            return [_act(np.concatenate([np.asarray(v) for v in x], axis=o["axis"]), o.get("act", 0))]
        if n == "SPLIT":
            axis = int(x[0])
            return list(np.split(x[1], o["num_splits"], axis=axis))
        if n == "GATHER":
            return [np.take(x[0], np.asarray(x[1]), axis=o.get("axis", 0))]
        if n == "SLICE":
            begin = [int(v) for v in x[1]]
            size = [int(v) for v in x[2]]
            idx = tuple(slice(b, None if s < 0 else b + s) for b, s in zip(begin, size))
            return [x[0][idx]]
        if n == "EXPAND_DIMS":
            return [np.expand_dims(x[0], int(x[1]))]
        if n == "CAST":
            return [x[0].astype(g.tensors[op.outputs[0]].dtype)]
        if n == "LESS":
            return [np.less(x[0], x[1])]
        if n == "PAD":
            return [np.pad(x[0], [(int(a), int(b)) for a, b in x[1]])]
        if n == "SPACE_TO_BATCH_ND":
            return [_space_to_batch(x[0], x[1], x[2])]
        if n == "BATCH_TO_SPACE_ND":
            return [_batch_to_space(x[0], x[1], x[2])]
        if n == "REDUCE_MAX":
            axes = tuple(int(v) for v in np.asarray(x[1]).reshape(-1))
            return [np.max(x[0], axis=axes, keepdims=bool(o.get("keep_dims", 0)))]
        if n == "WHILE":
            state = list(x)
            while bool(np.asarray(self._run(o["cond"], state)[0]).reshape(())):
                state = self._run(o["body"], state)
            return state
        raise NotImplementedError(n)
