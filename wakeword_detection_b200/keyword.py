"""`KeywordRecognizer` drop-in (reference: spokestack/asr/keyword/tflite.py:15-191).

Same constructor, stage protocol `recognizer(context, frame)`, `reset()` and `close()`: PCM-16 frames are scaled,
pre-emphasised (default 0.97) and framed; WHILE THE CONTEXT IS ACTIVE every completed 512-sample window becomes a mel
frame, the autoregressive encoder turns the mel window + its previous state into one encoded sample, and on the falling
edge of `context.is_active` the detector classifies the window of encoded samples (`recognize` / `timeout` events).

What runs where: the mel frames come from the fused CUDA filter kernel (csrc/filter.cu, fp64 FFT) - several frames per
call when a chunk completes several.  The reference ships NO keyword model (its tf_lite_models hold the two wake-word
families only), so there is no keyword encoder family in the CUDA library: the encoder / detector are the callables
given as `encode_model` / `detect_model` (anything with the reference TFLiteModel's call signature and
`input_details` / `output_details`); without them the constructor raises instead of guessing.
"""
from __future__ import annotations

import os
from typing import List

import numpy as np

from . import _cabi, weights as W


class KeywordRecognizer:
    def __init__(self, classes: List[str], pre_emphasis: float = 0.97, sample_rate: int = 16000,
                 fft_window_type: str = "hann", fft_hop_length: int = 10, model_dir: str = "",
                 posterior_threshold: float = 0.5, encode_model=None, detect_model=None, device: int = 0, **kwargs) -> None:
        self.classes = classes
        self.pre_emphasis: float = pre_emphasis
        self.hop_length: int = int(fft_hop_length * sample_rate / 1000)
        if fft_window_type != "hann":
            raise ValueError("Invalid fft_window_type")
        path, npz = os.path.join(model_dir, "filter.tflite"), os.path.join(model_dir, "weights.npz")
        key = (os.path.abspath(model_dir), "FILTER", int(device), "tc", False)
        eng = _cabi._ENGINES.get(key)
        if eng is None or eng.ctx is None:
            if os.path.isfile(path):
                fw = W.extract_filter(path)
            elif os.path.isfile(npz):          # the repository's extracted twin of a model directory
                with np.load(npz) as z:
                    fw = {k: z[k] for k in ("mel_w", "mel_b", "mel_floor", "mel_log_offset", "mel_scale")}
            else:
                raise ValueError("Could not open '%s'." % path)
            eng = _cabi.Engine(fw, device)
            _cabi._ENGINES[key] = eng
        self._engine = eng
        if encode_model is None or detect_model is None:
            raise ValueError("KeywordRecognizer: no keyword encoder family is built into the CUDA library (the reference ships no "
                             "keyword model); pass encode_model= and detect_model= callables")
        self.encode_model, self.detect_model = encode_model, detect_model
        if len(classes) != self.detect_model.output_details[0]["shape"][-1]:
            raise ValueError("Invalid number of classes")
        self._window_size = (eng.n_bins - 1) * 2
        if self._window_size != 512 or self.hop_length != 160:
            raise ValueError("the CUDA filter is built for a 512-sample window at a 160-sample hop")
        self.mel_length: int = int(self.encode_model.input_details[0]["shape"][1])
        self.mel_width: int = int(self.encode_model.input_details[0]["shape"][-1])
        if self.mel_width != eng.n_mel:
            raise ValueError("encoder expects %d mel bands, the filter produces %d" % (self.mel_width, eng.n_mel))
        self.state = np.zeros(self.encode_model.input_details[1]["shape"], np.float32)
        self.encode_length: int = int(self.detect_model.input_details[0]["shape"][1])
        self.encode_width: int = int(self.detect_model.input_details[0]["shape"][-1])
        self._posterior_threshold: float = posterior_threshold
        self._prev_sample: float = 0.0
        self._is_active = False
        self.reset()

    def __call__(self, context, frame) -> None:
        self._sample(context, frame)
        if not context.is_active and self._is_active:
            self._detect(context)
        self._is_active = context.is_active

    def _sample(self, context, frame) -> None:
        """:110-133: scale / clip, pre-emphasis with the carried sample, framing; windows completed while the context is
        inactive are consumed without analysis."""
        x = np.clip(np.asarray(frame).astype(np.float32) / (2 ** 15 - 1), -1.0, 1.0)
        if x.size == 0:
            return
        prev_sample = x[-1]
        x = x - self.pre_emphasis * np.append(self._prev_sample, x[:-1])      # float64 intermediate like numpy's
        self._prev_sample = prev_sample
        buf = np.concatenate([self._pending, x.astype(np.float32)])
        nf = self._engine.num_frames(buf.shape[0])
        if nf and context.is_active:
            mel = self._engine.filter(buf[None, :], 0.0)[0].cpu().numpy()
            for i in range(nf):
                self._encode(mel[i])
        self._pending = buf[nf * self.hop_length:] if nf else buf

    def _encode(self, mel_row: np.ndarray) -> None:
        """:152-173: push the mel frame, run the autoregressive encoder, push its output."""
        self._frames = np.concatenate([self._frames[1:], mel_row[None].astype(np.float32)])
        enc, self.state = self.encode_model(self._frames[None], self.state)
        self._encoded = np.concatenate([self._encoded[1:], np.asarray(enc, np.float32).reshape(1, self.encode_width)])

    def _detect(self, context) -> None:
        """:175-191."""
        posterior = self.detect_model(self._encoded[None])[0][0]
        class_index = int(np.argmax(posterior))
        confidence = posterior[class_index]
        if confidence >= self._posterior_threshold:
            context.transcript = self.classes[class_index]
            context.confidence = confidence
            context.event("recognize")
        else:
            context.event("timeout")
        self.reset()

    def reset(self) -> None:
        """:193-198 (the previous sample of the pre-emphasis filter is NOT reset there either)."""
        self._pending = np.zeros((0,), np.float32)
        self._frames = np.zeros((self.mel_length, self.mel_width), np.float32)        # frame_window.fill(0.0)
        self._encoded = np.full((self.encode_length, self.encode_width), -1.0, np.float32)   # encode_window.fill(-1.0)
        self.state[:] = 0.0

    def close(self) -> None:
        self.reset()
