"""Shared-column CRNN front vs the per-window path (WWB_CRNN_NO_SHARE=1): posteriors must be bit-identical."""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from conftest import get_engine
tc = get_engine("CRNN", "tc"); f32 = get_engine("CRNN", "f32")
ok = True
import sys as _s
SHAPES = [(1, 30000, 2), (3, 151 + 2 * 72, 2), (2, 151 + 127 * 2, 2), (2, 151 + 128 * 2, 2), (1, 9000, 1), (1000, 300, 2), (3, 5000, 4), (2, 9000, 8), (149, 998, 2), (2, 1015, 2), (2, 1015, 1), (2, 1017, 2), (2, 1023, 2), (3, 2023, 2)] if len(_s.argv) > 1 and _s.argv[1] == 'edge' else None
for (S, F, hop) in SHAPES or [(3, 200, 2), (2, 151 + 2 * 130, 2), (5, 998, 2), (2, 700, 1), (3, 600, 4), (2, 1200, 8), (1, 153, 2), (7, 2000, 2), (90, 998, 2)]:
    torch.manual_seed(S * 1000 + F)
    X = torch.rand((S, F, 40), device=tc.device) * 5
    os.environ["WWB_CRNN_NO_SHARE"] = "1"
    a = tc.posteriors(X, hop=hop).clone(); torch.cuda.synchronize()
    os.environ["WWB_CRNN_NO_SHARE"] = "0"
    b = tc.posteriors(X, hop=hop).clone(); torch.cuda.synchronize()
    b2 = tc.posteriors(X, hop=hop).clone(); torch.cuda.synchronize()
    d = (a - b).abs().max().item()
    ref = f32.posteriors(X, hop=hop)
    e = (b - ref).abs().max().item()
    ok &= e < 1e-4
    print("   vs fp32 path: max err %.3e" % e)
    same = bool((a == b).all().item()) and bool((b == b2).all().item())
    ok &= same
    print("S=%d F=%d hop=%d n_win=%d: max |diff| %.3e bit-identical=%s" % (S, F, hop, a.shape[1], d, same), flush=True)
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
