"""Minimal stand-in for the reference's SpeechContext (reference:
spokestack/context.py:12-128): only the two flags the wake-word stage reads and
writes (`is_speech`, `is_active`).  Any object with these attributes works — the real
spokestack SpeechContext included; the pipeline runtime itself is out of scope."""


class SpeechContext:
    def __init__(self) -> None:
        self.is_speech: bool = False
        self.is_active: bool = False
        self.events = []

    def event(self, name: str) -> None:
        self.events.append(name)
