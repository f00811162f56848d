"""`TFLiteModel`-compatible callables backed by the CUDA library.

Drop-in for the reference's model runner (reference: spokestack/models/tensorflow.py:15-69
and its three byte-identical copies): `TFLiteModel(model_path)(*arrays) -> [np.ndarray]`,
`.input_details[0]["shape"]`, `.output_details`.  The file name selects the role
(filter / encode / detect, as every reference call site names them); the three models
of a directory share one device context.  Unlike the interpreter, the leading batch
dimension may be any B >= 1.
"""
from __future__ import annotations

import os
from typing import Any, List

import numpy as np

from . import _cabi, weights as W


def _kind_of_dir(model_dir: str) -> str:
    npz = os.path.join(model_dir, "weights.npz")
    enc = os.path.join(model_dir, "encode.tflite")
    if os.path.isfile(enc):
        from . import tflite_reader as tr
        shape = [int(v) for v in tr.load(enc).input_details()[0]["shape"]]
        return "CRNN" if len(shape) == 4 else "Wavenet"
    if os.path.isfile(npz):
        with np.load(npz) as z:
            return "CRNN" if "conv_w" in z.files else "Wavenet"
    return ""


class TFLiteModel:
    """model_path: <dir>/filter.tflite | <dir>/encode.tflite | <dir>/detect.tflite, or the float16 variants
    <dir>/encode-quant.tflite | <dir>/detect-quant.tflite (reference: utils/evaluate_tf_lite_opts.py:16-33).  A
    `-quant` file that does not exist raises like the interpreter would, unless `derive_quant=True`: then the variant is
    derived from the float32 files by the converter's rounding (weights.quantize_fp16).
    `device` / `precision` / `derive_quant` are extensions (keyword only)."""

    def __init__(self, model_path: str, **kwargs: Any) -> None:
        self.model_path = model_path
        model_dir = os.path.dirname(model_path) or "."
        stem = os.path.basename(model_path).split(".")[0]
        role, _, variant = stem.partition("-")
        if role not in ("filter", "encode", "detect") or variant not in ("", "quant") or (role == "filter" and variant):
            raise ValueError("Could not open '%s': expected filter/encode/detect[-quant] .tflite" % model_path)
        quant = variant == "quant"
        derive = bool(kwargs.pop("derive_quant", False))
        have_file = os.path.isfile(model_path) or (not quant and os.path.isfile(os.path.join(model_dir, "weights.npz")))
        if quant and not have_file and derive:
            have_file = os.path.isfile(os.path.join(model_dir, role + ".tflite")) or os.path.isfile(os.path.join(model_dir, "weights.npz"))
        if not have_file:
            raise ValueError("Could not open '%s'." % model_path)
        self.role = role
        self.quant = quant
        device = int(kwargs.pop("device", 0))
        precision = kwargs.pop("precision", "tc")
        kind = _kind_of_dir(model_dir)
        if role != "filter" and not kind:
            raise ValueError("Could not open '%s': no encode/detect weights in %s" % (model_path, model_dir))
        if kind:
            self._engine = _cabi.engine_for_dir(model_dir, kind, device, precision, quant=quant)
        else:
            key = (os.path.abspath(model_dir), "FILTER", device, precision, False)
            eng = _cabi._ENGINES.get(key)
            if eng is None or eng.ctx is None:
                eng = _cabi.Engine(W.extract_filter(os.path.join(model_dir, "filter.tflite")), device, precision)
                _cabi._ENGINES[key] = eng
            self._engine = eng
        e = self._engine
        crnn = e.kind == _cabi.WWB_MODEL_CRNN
        if role == "filter":
            ins, outs = [1, e.n_bins], [1, e.n_mel]
        elif role == "encode":
            ins = [1, e.n_mel, e.L, 1] if crnn else [1, e.L, e.n_mel]
            outs = [1, 64] if crnn else [1, e.L, 32]
        else:
            ins = [1, 64] if crnn else [1, e.L, 32]
            outs = [1, e.n_out]
        self._input_details = [{"name": role + "_input", "index": 0, "shape": np.array(ins, np.int32),
                                "dtype": np.float32}]
        self._output_details = [{"name": role + "_output", "index": 1, "shape": np.array(outs, np.int32),
                                 "dtype": np.float32}]

    def __call__(self, *args) -> List[np.ndarray]:
        if len(args) != 1:
            raise ValueError("expected exactly one input tensor")
        return [self.run_device(args[0]).cpu().numpy()]

    def run_device(self, x):
        """Same as __call__ but returns the torch CUDA tensor (no device->host copy)."""
        e = self._engine
        torch = e.torch
        if not isinstance(x, torch.Tensor):
            x = np.asarray(x)
            if x.dtype != np.float32:
                raise ValueError("Cannot set tensor: Got value of type %s but expected type FLOAT32" % x.dtype)
        want = list(self._input_details[0]["shape"][1:])
        if list(x.shape[1:]) != want or len(x.shape) != len(want) + 1:
            raise ValueError("Cannot set tensor: Dimension mismatch. Got %s but expected [B, %s]"
                             % (list(x.shape), ", ".join(map(str, want))))
        if self.role == "filter":
            return e.mel_from_magnitude(x)
        if self.role == "encode":
            if e.kind == _cabi.WWB_MODEL_CRNN:
                x = e._dev(x, torch.float32)[..., 0].transpose(1, 2)   # [B,40,151,1] -> [B,151,40]
            return e.encode(x)
        return e.detect(x)

    @property
    def input_details(self) -> List[Any]:
        return self._input_details

    @property
    def output_details(self) -> List[Any]:
        return self._output_details
