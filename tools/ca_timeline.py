# needs a library built with NVCC_EXTRA=-DWWB_TIMELINE python -m wakeword_detection_b200.build --force
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wakeword_detection_b200 import _cabi, weights as W
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = _cabi.Engine(W.load_model_dir(os.path.join(ROOT, "weights", "CRNN"), "CRNN"), 0, sys.argv[1] if len(sys.argv) > 1 else "tc")
mel = torch.rand((148 * 6 * 4, 151, 40), device=eng.device) * 5
eng.posteriors(mel, hop=1); torch.cuda.synchronize()
dbg = torch.zeros(1024, dtype=torch.int64, device=eng.device)
eng.lib.wwb_debug_buffer(eng.ctx, dbg.data_ptr())
eng.posteriors(mel, hop=1); torch.cuda.synchronize()
d = dbg.cpu().numpy()[:160].reshape(20, 8)
t0 = d[0, 0]
np.set_printoptions(linewidth=220)
print("per f: [xp_wait_start, xp_done, cacc_wait_done, conv_issued | inproj(f): start, a1_full, w1_full, issued]  (relative clk)")
for f in range(20):
    r = d[f] - t0
    print(f, r, " xpwait=%d caccwait=%d convissue=%d | a1wait=%d w1wait=%d projissue=%d" % (r[1]-r[0], r[2]-r[1], r[3]-r[2], r[5]-r[4], r[6]-r[5], r[7]-r[6]))
print("period per f:", np.diff(d[:, 0]))
