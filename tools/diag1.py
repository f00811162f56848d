import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import get_engine, load_weights
from oracle import restated as R
from wakeword_detection_b200 import synth
from test_gpu_parity import _windows
w = load_weights("CRNN"); eng = get_engine("CRNN")
for cls in range(6):
    pcm = synth.stream_int16(16000 + 37 * cls, cls, 1, cls)
    mel = eng.filter(pcm[None]).cpu().numpy()[0].astype(np.float64)
    ref = R.mel_stream(R.int16_to_float(pcm), w).astype(np.float64)
    err = np.abs(mel - ref)
    for lo, hi in ((0, 0.35), (0.35, 1.15), (1.15, 2.3), (2.3, 99)):
        m = (ref >= lo) & (ref < hi)
        if m.any():
            print("cls %d ref in [%.2f,%.2f): n=%d max err %.2e  max rel(max(ref,1)) %.2e" % (cls, lo, hi, m.sum(), err[m].max(), (err[m] / np.maximum(ref[m], 1)).max()))
for wname in ("CRNN", "CRNN_arik_original"):
    w = load_weights(wname); eng = get_engine(wname)
    X = _windows("crnn", w)
    enc = eng.encode(X).cpu().numpy(); ref_enc = R.encode(X, w)
    print(wname, "enc err", np.abs(enc - ref_enc).max())
    det = eng.detect(enc).cpu().numpy(); ref_det = R.detect(ref_enc, w)
    print(wname, "det err", np.abs(det - ref_det).max(), det[:3], ref_det[:3])
    post = eng.posteriors(X, hop=1).cpu().numpy()[:, 0]
    print(wname, "post err", np.abs(post - ref_det[:, -1]).max(), "range", ref_det[:, -1].min(), ref_det[:, -1].max())
