"""Race hunt: the encoders are deterministic, so repeated runs on the same input must be bit-identical.
    python tools/stress_determinism.py [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from wakeword_detection_b200 import _cabi, weights as W, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for name in ("Wavenet", "CRNN"):
    eng = _cabi.Engine(W.load_model_dir(os.path.join(ROOT, "weights", name), name), 0, "tc")
    for S, sec in ((512, 10.0), (37, 3.3), (1, 2.0)):
        pcm = synth.device_pcm(S, int(sec * 16000), seed=7 + S, device=eng.device)
        mel = eng.filter(pcm, 0.0)
        ref = eng.posteriors(mel, 2).clone()
        bad = 0
        for i in range(reps):
            out = eng.posteriors(mel, 2)
            if not torch.equal(out, ref):
                bad += 1
                print("  MISMATCH run %d: max |diff| %.3e, %d elements" % (i, float((out - ref).abs().max()), int((out != ref).sum())))
        print("%s S=%d %.1fs windows=%d: %d/%d runs differ" % (name, S, sec, ref.numel(), bad, reps))
    eng.close()
