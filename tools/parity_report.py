"""Measured parity of the CUDA path against the numpy oracle (max abs error), for DESIGN.md §4.
    python tools/parity_report.py        (needs a B200; prints one line per model / precision)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import get_engine, load_weights
from oracle import restated as R
from test_gpu_parity import _windows

for wname, name in (("CRNN", "crnn"), ("CRNN_arik_original", "crnn"), ("Wavenet", "wavenet")):
    w = load_weights(wname)
    X = _windows(name, w)
    if name == "wavenet":
        X = X[::2]
    ref_enc = R.encode(X, w)
    ref = R.detect(ref_enc, w)[:, -1]
    for prec in ("f32", "tc", "tc_fast"):
        eng = get_engine(wname, prec)
        enc = eng.encode(X).cpu().numpy()
        post = eng.posteriors(X.reshape(X.shape[0], X.shape[1], 40), hop=1).cpu().numpy()[:, 0]
        band = np.abs(ref - 0.5) <= 1e-3
        flips = int(((post > 0.5) != (ref > 0.5))[~band].sum())
        print("%-20s %-8s windows %4d  max |enc err| %.2e  max |posterior err| %.2e  decision flips outside the band %d"
              % (wname, prec, X.shape[0], np.abs(enc - ref_enc).max(), np.abs(post - ref).max(), flips))

# sliding windows (hop 2, get_posterior semantics): the CRNN tensor-core path shares conv / GRU-1 projection columns
# between the overlapping windows of a stream here (crnn_tc.cu, CrnnShare)
from wakeword_detection_b200 import synth
for wname in ("CRNN", "Wavenet"):
    w = load_weights(wname)
    L = int(w["mel_length"])
    mels = np.stack([R.mel_stream(np.clip(synth.stream_float(16000 * 6, c, 5, c), -1, 1).astype(np.float32), w)
                     for c in range(synth.N_CLASSES)])
    nw = R.eval_windows(mels.shape[1], L)
    j = np.arange(0, nw, 7)
    ref = np.stack([R.posterior(m[(2 * j)[:, None] + np.arange(L)[None, :]], w) for m in mels])
    for prec in ("f32", "tc"):
        post = get_engine(wname, prec).posteriors(mels, hop=2).cpu().numpy()[:, j]
        print("%-20s %-8s sliding hop-2 windows %d streams x %d (every 7th checked)  max |posterior err| %.2e"
              % (wname, prec, mels.shape[0], nw, np.abs(post - ref).max()))
