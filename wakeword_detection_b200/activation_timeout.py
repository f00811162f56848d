"""`ActivationTimeout` drop-in (reference: spokestack/activation_timeout.py:7-50): the pipeline stage that ends an
activation after `min_active` ms on a VAD fall, or after `max_active` ms.  Host-side, one stream; the many-stream form
runs on the device next to the trigger (`wakeword.MultiStreamPipeline`, csrc/context.cu)."""
from __future__ import annotations


class ActivationTimeout:
    """Speech pipeline activation timeout: frame_width / min_active / max_active in ms."""

    def __init__(self, frame_width: int = 20, min_active: int = 500, max_active: int = 5000, **kwargs) -> None:
        self._min_active = min_active / frame_width
        self._max_active = max_active / frame_width
        self._is_speech = False
        self._active_length = 0

    def __call__(self, context, frame=None) -> None:
        fell = self._is_speech and not context.is_speech
        self._is_speech = context.is_speech
        if not context.is_active:
            return
        self._active_length += 1
        if self._active_length > self._min_active and (fell or self._active_length > self._max_active):
            self.deactivate(context)

    def deactivate(self, context) -> None:
        self.reset()
        context.is_active = False

    def reset(self) -> None:
        self.close()

    def close(self) -> None:
        self._active_length = 0
