"""Timeline of the two-phase WaveNet schedule (needs a build with NVCC_EXTRA=-DWWB_WN_TIMELINE).
WWB_WN_DBG_EARLY=1 stamps the early phase (blocks 0..cut-1, 8 windows per group), otherwise the late phase."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wakeword_detection_b200 import _cabi, weights as W
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = _cabi.Engine(W.load_model_dir(os.path.join(ROOT, "weights", "Wavenet"), "Wavenet"), 0, "tc")
mel = torch.rand((64, 998, 40), device=eng.device) * 5
eng.posteriors(mel, hop=2); torch.cuda.synchronize()
dbg = torch.zeros(8 * 48 * 4 + 64 + 24 * 16, dtype=torch.int64, device=eng.device)
eng.lib.wwb_debug_buffer(eng.ctx, dbg.data_ptr())
eng.posteriors(mel, hop=2); torch.cuda.synchronize()
d = dbg.cpu().numpy()[:8 * 48 * 4].reshape(8, 48, 4)
t0 = d[5][d[5] > 0].min()
np.set_printoptions(linewidth=220)
g = d[5, :, 0]
print("gate issue of tile 0 per block, two consecutive groups (clk):", (g[g > 0] - t0).tolist())
print("period per block:", np.diff(g[g > 0]).tolist())
for k in range(24):
    rows = []
    for tile in range(5):
        e = d[tile, k]
        if e[0] > 0:
            rows.append("t%d e1=%d rs=%d e2a=%d" % (tile, e[1] - e[0], e[2] - e[1], e[3] - e[2]))
    if rows:
        print("block %2d: %s" % (k, " | ".join(rows)))
print("boundary (tile 0): ", (d[6, 20, :4] - t0).tolist(), (d[6, 21, :2] - t0).tolist())
