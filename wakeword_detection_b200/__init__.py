"""B200-native filter -> encode -> detect wake-word hot path (drop-in for the
TFLite interpreter calls of MerlinPCarson/WakeWord-Detection)."""
__version__ = "0.1.0"
