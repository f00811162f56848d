"""`Filter` — offline featuriser drop-in (reference: utils/tf_lite/filter.py:7-79).

Same constructor and `filter_frame(frame) -> list[np.ndarray(40)]` / `num_outputs()`
behaviour: samples accumulate in a 512-sample window that advances by the hop, state
(window content, previous sample for pre-emphasis) carries across calls, and the
caller's `frame` is pre-emphasised in place like the reference does (:43).  The
arithmetic runs in the fused CUDA filter kernel (csrc/filter.cu); `filter_streams` is
the batched form the evaluation path uses.
"""
from __future__ import annotations

import os
from typing import List

import numpy as np

from . import _cabi, weights as W

FFT = 512


class Filter:
    def __init__(self, pre_emphasis: float = 0.0, sample_rate: int = 16000,
                 fft_window_type: str = "hann", fft_hop_length: int = 10, model_dir: str = "",
                 device: int = 0, engine: "_cabi.Engine" = None) -> None:
        self.pre_emphasis: float = pre_emphasis
        self.hop_length: int = int(fft_hop_length * sample_rate / 1000)
        if fft_window_type != "hann":
            raise ValueError("Invalid fft_window_type")
        if engine is None:
            from .models import _kind_of_dir
            kind = _kind_of_dir(model_dir)
            if kind:
                engine = _cabi.engine_for_dir(model_dir, kind, device)
            else:
                npz = os.path.join(model_dir, "weights.npz")
                path = os.path.join(model_dir, "filter.tflite")
                if os.path.isfile(path):
                    fw = W.extract_filter(path)
                elif os.path.isfile(npz):
                    with np.load(npz) as z:
                        fw = {k: z[k] for k in z.files}
                else:
                    raise ValueError("Could not open '%s'." % path)
                engine = _cabi.Engine(fw, device)
        self._engine = engine
        self._window_size = (engine.n_bins - 1) * 2
        if self._window_size != FFT or self.hop_length != 160:
            raise ValueError("the CUDA filter is built for a 512-sample window at a 160-sample hop")
        self._pending = np.zeros((0,), np.float32)   # unread content of the sample window
        self._prev_sample: float = 0.0

    # reference: filter.py:38-57
    def filter_frame(self, frame) -> List[np.ndarray]:
        frame = np.asarray(frame) if not isinstance(frame, np.ndarray) else frame
        if frame.size == 0:
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")
        prev_sample = frame[-1]
        frame -= self.pre_emphasis * np.append(self._prev_sample, frame[:-1])
        self._prev_sample = prev_sample
        buf = np.concatenate([self._pending, frame.astype(np.float32, copy=False)])
        nf = self._engine.num_frames(buf.shape[0])
        if nf == 0:
            self._pending = buf
            return []
        mel = self._engine.filter(buf[None, :], 0.0)[0].cpu().numpy()
        self._pending = buf[nf * self.hop_length:]
        return [mel[i] for i in range(nf)]

    def num_outputs(self) -> int:
        return int(self._engine.n_mel)

    # batched extension -----------------------------------------------------------------
    def filter_streams(self, pcm, pre_emphasis: float = None):
        """[S, N] int16|float32 (numpy or torch) -> torch CUDA tensor [S, F, 40]; every row
        is an independent stream starting from an empty window."""
        a = self.pre_emphasis if pre_emphasis is None else pre_emphasis
        return self._engine.filter(pcm, a)
