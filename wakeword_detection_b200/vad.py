"""`VoiceActivityDetector` / `VoiceActivityTrigger` drop-ins (reference: spokestack/vad/webrtc.py:21-126).

The per-frame speech decision itself comes from the webrtcvad C extension in the reference; it is outside the
filter -> encode -> detect path and not installable here, so the decision function is injectable:
`detector(frame_bytes, sample_rate) -> bool` (webrtcvad.Vad(mode).is_speech when the package is present).  What this
module re-implements is the part the trigger depends on: the run-length rise / fall debounce that drives
`context.is_speech` (:62-77).  The many-stream form runs on the device (csrc/context.cu)."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

QUALITY = 0
LOW_BITRATE = 1
AGGRESSIVE = 2
VERY_AGGRESSIVE = 3


class VoiceActivityDetector:
    def __init__(self, sample_rate: int = 16000, frame_width: int = 20, vad_rise_delay: int = 0, vad_fall_delay: int = 0,
                 mode: int = QUALITY, detector: Optional[Callable[[bytes, int], bool]] = None, **kwargs) -> None:
        self._sample_rate = sample_rate
        self._rise_length = vad_rise_delay // frame_width
        self._fall_length = vad_fall_delay // frame_width
        if detector is None:
            try:
                import webrtcvad  # type: ignore
            except Exception as e:
                raise ImportError("webrtcvad is not installed: pass detector=callable(frame_bytes, sample_rate)") from e
            detector = webrtcvad.Vad(mode).is_speech
        self._detect = detector
        self._run_value = 0
        self._run_length = 0

    def __call__(self, context, frame: np.ndarray) -> None:
        raw = bool(self._detect(np.asarray(frame).tobytes(), self._sample_rate))
        if raw == self._run_value:
            self._run_length += 1
        else:
            self._run_value, self._run_length = raw, 1
        if self._run_value != context.is_speech:
            if self._run_value and self._run_length >= self._rise_length:
                context.is_speech = True
            if not self._run_value and self._run_length >= self._fall_length:
                context.is_speech = False

    def reset(self) -> None:
        self._run_value = 0
        self._run_length = 0

    def close(self) -> None:
        self.reset()


class VoiceActivityTrigger:
    """Activates the context on a VAD rise (:103-126)."""

    def __init__(self) -> None:
        self._is_speech = False

    def __call__(self, context, frame=None) -> None:
        if context.is_speech != self._is_speech:
            if context.is_speech:
                context.is_active = True
            self._is_speech = context.is_speech

    def close(self) -> None:
        self.reset()

    def reset(self) -> None:
        self._is_speech = False
