"""Deterministic stand-ins for a keyword model pair (the reference ships no keyword models): an autoregressive encoder
`encode(frames [1, 12, 40], state [1, 8]) -> (enc [1, 6], state [1, 8])` and a detector
`detect(encs [1, 5, 6]) -> [1, 3]` (softmax), with the `input_details` / `output_details` the reference's
KeywordRecognizer reads its geometry from (spokestack/asr/keyword/tflite.py:57-79).  Used by
make_golden_keyword.py (under the reference's class) and by the tests (under the drop-in and the oracle)."""
import numpy as np

MEL_LENGTH, MEL_WIDTH, STATE, ENC_LENGTH, ENC_WIDTH, CLASSES = 12, 40, 8, 5, 6, 3
_rng = np.random.default_rng(123)
_WS = (_rng.standard_normal((MEL_WIDTH, STATE)) * 0.05).astype(np.float32)
_WE = (_rng.standard_normal((STATE, ENC_WIDTH)) * 0.8).astype(np.float32)
_WD = (_rng.standard_normal((ENC_WIDTH, CLASSES)) * 2.0).astype(np.float32)


def _d(shape):
    return [{"shape": np.array(s, np.int32), "dtype": np.float32, "index": i} for i, s in enumerate(shape)]


class Encode:
    input_details = _d([[1, MEL_LENGTH, MEL_WIDTH], [1, STATE]])
    output_details = _d([[1, ENC_WIDTH], [1, STATE]])

    def __call__(self, frames, state):
        frames, state = np.asarray(frames, np.float32), np.asarray(state, np.float32)
        assert frames.shape == (1, MEL_LENGTH, MEL_WIDTH) and state.shape == (1, STATE)
        new = np.tanh(np.float32(0.8) * state + frames.mean(axis=1) @ _WS).astype(np.float32)
        return [(new @ _WE).astype(np.float32), new]


class Detect:
    input_details = _d([[1, ENC_LENGTH, ENC_WIDTH]])
    output_details = _d([[1, CLASSES]])

    def __call__(self, encs):
        encs = np.asarray(encs, np.float32)
        assert encs.shape == (1, ENC_LENGTH, ENC_WIDTH)
        z = encs.mean(axis=1) @ _WD
        e = np.exp(z - z.max(axis=1, keepdims=True))
        return [(e / e.sum(axis=1, keepdims=True)).astype(np.float32)]
