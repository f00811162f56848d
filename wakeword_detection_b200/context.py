"""Minimal stand-in for the reference's SpeechContext (reference:
spokestack/context.py:12-128): the two flags the wake-word stage reads and writes (`is_speech`, `is_active`), the
transcript / confidence the keyword recognizer sets, and named event handlers.  Any object with these attributes works -
the real spokestack SpeechContext included; the pipeline runtime itself is out of scope."""


class SpeechContext:
    def __init__(self) -> None:
        self.is_speech: bool = False
        self.is_active: bool = False
        self.transcript: str = ""
        self.confidence: float = 0.0
        self.events = []
        self._handlers = {}

    def add_handler(self, name: str, function) -> None:
        self._handlers[name] = function

    def event(self, name: str) -> None:
        self.events.append(name)
        handler = self._handlers.get(name)
        if handler:
            handler(self)
