"""Golden run of the reference's OWN KeywordRecognizer (spokestack/asr/keyword/tflite.py:15-191) with
filter.tflite executed by the literal interpreter and the stand-in keyword model pair of keyword_stub.py
(the reference ships no keyword models).  Pins the glue: int16 scaling, pre-emphasis 0.97, analysis only while the
context is active, autoregressive encoder state, encode window pre-filled with -1, classification on the falling edge
of is_active, reset.

    python tests/golden/make_golden_keyword.py      # writes tests/golden/reference_keyword.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import keyword_stub as KS  # noqa: E402
import make_golden as MG  # noqa: E402
from wakeword_detection_b200 import synth  # noqa: E402


def main():
    MG.install_stubs()
    sys.path.insert(0, REF)
    import spokestack.models.tensorflow as TF
    real = TF.TFLiteModel

    def factory(model_path, **kw):
        name = os.path.basename(model_path)
        if name == "encode.tflite":
            return KS.Encode()
        if name == "detect.tflite":
            return KS.Detect()
        return real(model_path=model_path)

    import spokestack.asr.keyword.tflite as KW
    KW.TFLiteModel = factory
    from spokestack.context import SpeechContext

    classes = ["up", "down", "stop"]
    rec = KW.KeywordRecognizer(classes=classes, model_dir=os.path.join(REF, "tf_lite_models/CRNN"), posterior_threshold=0.9)
    ctx = SpeechContext()
    events = []
    ctx.add_handler("recognize", lambda c: events.append(("recognize", c.transcript, float(c.confidence))))
    ctx.add_handler("timeout", lambda c: events.append(("timeout", "", 0.0)))
    pcm = np.concatenate([synth.stream_int16(16000, c, 21, c) for c in (2, 0, 5, 2, 4)])
    n_frames = pcm.shape[0] // 320
    active = np.zeros(n_frames, bool)
    active[10:60] = True
    active[100:103] = True          # a very short activation: the encode window is still mostly -1
    active[150:230] = True
    enc_log, ev_frame = [], []
    for i in range(n_frames):
        ctx.is_active = bool(active[i])
        n0 = len(events)
        rec(ctx, pcm[i * 320:(i + 1) * 320].copy())
        if len(events) > n0:
            ev_frame.append(i)
        enc_log.append(rec.encode_window.read_all().copy() if not rec.encode_window.is_empty else np.zeros((KS.ENC_LENGTH, KS.ENC_WIDTH), np.float32))
        rec.encode_window.rewind()
    out = {"kw_pcm": pcm[:n_frames * 320], "kw_active": active, "kw_event_frame": np.array(ev_frame, np.int32),
           "kw_event_kind": np.array([e[0] for e in events]), "kw_event_class": np.array([e[1] for e in events]),
           "kw_event_conf": np.array([e[2] for e in events], np.float32),
           "kw_enc_window": np.stack(enc_log).astype(np.float32), "kw_threshold": np.array(0.9, np.float32)}
    np.savez_compressed(os.path.join(HERE, "reference_keyword.npz"), **out)
    print(events, ev_frame)


if __name__ == "__main__":
    main()
