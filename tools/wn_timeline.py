# needs a library built with NVCC_EXTRA=-DWWB_TIMELINE python -m wakeword_detection_b200.build --force
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wakeword_detection_b200 import _cabi, weights as W
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = _cabi.Engine(W.load_model_dir(os.path.join(ROOT, "weights", "Wavenet"), "Wavenet"), 0, sys.argv[1] if len(sys.argv) > 1 else "tc")
mel = torch.rand((148 * 3 * 4, 182, 40), device=eng.device) * 5
eng.posteriors(mel, hop=1); torch.cuda.synchronize()
dbg = torch.zeros(8 * 48 * 4, dtype=torch.int64, device=eng.device)
eng.lib.wwb_debug_buffer(eng.ctx, dbg.data_ptr())
eng.posteriors(mel, hop=1); torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(8, 48, 4)
t0 = d[5, 0, 0]
np.set_printoptions(linewidth=200)
for k in range(2, 7):
    print("block", k, "issuer G0,G4,R0,R4:", d[5, k] - t0, " G1 wait_start,wait_end,issued:", d[6, k, :3] - t0, " R3 wait_start,wait_end,issued (block k):", d[7, k, :3] - t0)
    for tile in range(5):
        print("   tile", tile, "gate_wake, e1_done, rs_wake, e2_done:", d[tile, k] - t0, " e1=%d e2=%d" % (d[tile, k, 1] - d[tile, k, 0], d[tile, k, 3] - d[tile, k, 2]))
print("period per block (issuer G0), two consecutive groups:", np.diff(d[5, :, 0]))
print("tile 0..4 gate_wake around the group boundary (blocks 22, 23 | 0, 1):", [(d[t, 22:26, 0] - t0).tolist() for t in range(5)])
print("tile 0..4 e2a done (blocks 22, 23 | 0, 1):", [(d[t, 22:26, 3] - t0).tolist() for t in range(5)])
print("tile 0 boundary (rel. clk): e2b(23) done, detect GEMM done, detect epilogue done, barrier+finalise done | next group: mel chunks stored, input GEMM done, first gate_wake:",
      (d[6, 20, :4] - t0).tolist(), (d[6, 24 + 21, :2] - t0).tolist(), int(d[0, 24, 0] - t0))
