"""Golden run of the reference's OWN pipeline stages around the wake-word trigger:
spokestack/vad/webrtc.py VoiceActivityDetector (rise / fall debounce) -> spokestack/wakeword/tflite.py WakewordTrigger
-> spokestack/activation_timeout.py ActivationTimeout, dispatched frame by frame like SpeechPipeline._dispatch
(spokestack/pipeline.py:25-28).  Third-party pieces are stubbed as in make_golden.py (TFLite -> the literal
interpreter); webrtcvad.Vad.is_speech returns a scripted raw decision per frame (the C extension is not installed and
the raw decision is an INPUT of the path: SURVEY.md 2 marks the VAD itself out of scope).

    python tests/golden/make_golden_pipeline.py      # writes tests/golden/reference_pipeline.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as MG  # noqa: E402
from wakeword_detection_b200 import synth  # noqa: E402

SCRIPT = {"raw": None, "i": 0}


def main():
    MG.install_stubs()
    vad_mod = types.ModuleType("webrtcvad")

    class Vad:
        def __init__(self, mode=0):
            self.mode = mode

        def is_speech(self, frame_bytes, sample_rate):
            v = bool(SCRIPT["raw"][SCRIPT["i"]])
            SCRIPT["i"] += 1
            return v

    vad_mod.Vad = Vad
    sys.modules["webrtcvad"] = vad_mod
    sys.path.insert(0, REF)
    os.chdir(REF)
    from spokestack.wakeword.tflite import WakewordTrigger
    from spokestack.context import SpeechContext
    from spokestack.activation_timeout import ActivationTimeout
    from spokestack.vad.webrtc import VoiceActivityDetector

    out = {}
    cfg = dict(frame_width=20, vad_rise_delay=40, vad_fall_delay=60, min_active=200, max_active=1000)
    out["cfg"] = np.array([cfg[k] for k in ("frame_width", "vad_rise_delay", "vad_fall_delay", "min_active", "max_active")], np.int32)
    for name, sub, typ in (("crnn", "CRNN", "CRNN"), ("wavenet", "Wavenet", "Wavenet")):
        wake = np.load(os.path.join(HERE, "wake_%s_pcm.npy" % name))
        pcm = np.concatenate([synth.stream_int16(3200, 0, 6, 1), wake, synth.stream_int16(16000, 1, 6, 2), wake[:24000],
                              synth.stream_int16(8000, 2, 6, 3)])
        n_frames = pcm.shape[0] // 320
        rng = np.random.default_rng(3)
        raw = np.ones(n_frames, bool)
        raw[:5] = False
        raw[9] = False                                  # one-frame glitch: shorter than the fall delay
        raw[60:62] = False                              # two-frame gap: still shorter
        raw[118:140] = False                            # a real pause: VAD falls, wake windows reset
        raw[141] = False
        raw[n_frames - 30:] = rng.random(30) < 0.5      # chatter at the end
        SCRIPT["raw"], SCRIPT["i"] = raw, 0
        vad = VoiceActivityDetector(sample_rate=16000, **cfg)
        trig = WakewordTrigger(model_dir=os.path.join(REF, "tf_lite_models", sub), model_type=typ)
        tmo = ActivationTimeout(**cfg)
        ctx = SpeechContext()
        speech, active, pmax = [], [], []
        for i in range(n_frames):
            frame = pcm[i * 320:(i + 1) * 320]
            for stage in (vad, trig, tmo):              # SpeechPipeline._dispatch
                stage(ctx, frame)
            speech.append(ctx.is_speech)
            active.append(ctx.is_active)
            pmax.append(trig._posterior_max)
        out["pipe_%s_pcm" % name] = pcm[:n_frames * 320]
        out["pipe_%s_raw" % name] = raw
        out["pipe_%s_speech" % name] = np.array(speech, bool)
        out["pipe_%s_active" % name] = np.array(active, bool)
        out["pipe_%s_post_max" % name] = np.array(pmax, np.float32)
        a = np.array(active, bool)
        print(name, "frames", n_frames, "speech frames", int(np.sum(speech)), "active frames", int(a.sum()),
              "activations at", np.nonzero(a & ~np.concatenate([[False], a[:-1]]))[0].tolist(),
              "deactivations at", np.nonzero(~a & np.concatenate([[False], a[:-1]]))[0].tolist())
    np.savez_compressed(os.path.join(HERE, "reference_pipeline.npz"), **out)


if __name__ == "__main__":
    main()
