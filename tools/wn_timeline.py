# needs a library built with -DWWB_TIMELINE: tools/build_variant.sh tl wavenet_tc.cu -DWWB_TIMELINE ; WWB200_LIB=build/libwwb200_tl.so
# (per-window schedule: independent windows, all tiles from block 0; roles of the debug buffer: tiles 0..NT-1, issuers, boundary, res/skip warp)
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wakeword_detection_b200 import _cabi, weights as W
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = _cabi.Engine(W.load_model_dir(os.path.join(ROOT, "weights", "Wavenet"), "Wavenet"), 0, sys.argv[1] if len(sys.argv) > 1 else "tc")
NT = 6
NR = NT + 3
mel = torch.rand((148 * 4 * 4, 182, 40), device=eng.device) * 5
eng.posteriors(mel, hop=1); torch.cuda.synchronize()
dbg = torch.zeros(NR * 48 * 4 + 64 + 24 * 16, dtype=torch.int64, device=eng.device)
eng.lib.wwb_debug_buffer(eng.ctx, dbg.data_ptr())
eng.posteriors(mel, hop=1); torch.cuda.synchronize()
d = dbg.cpu().numpy()[:NR * 48 * 4].reshape(NR, 48, 4)
t0 = d[NT, 0, 0]
np.set_printoptions(linewidth=200)
for k in range(2, 7):
    print("block", k, "issuer G0,G4,R0,R4:", d[NT, k] - t0, " G1 wait_start,wait_end,issued:", d[NT + 1, k, :3] - t0, " R3 wait_start,wait_end,issued (block k):", d[NT + 2, k, :3] - t0)
    for tile in range(NT):
        print("   tile", tile, "gate_wake, e1_done, rs_wake, e2_done:", d[tile, k] - t0, " e1=%d e2=%d" % (d[tile, k, 1] - d[tile, k, 0], d[tile, k, 3] - d[tile, k, 2]))
print("period per block (issuer G0), two consecutive groups:", np.diff(d[NT, :, 0]))
print("tile 0..5 gate_wake around the group boundary (blocks 22, 23 | 0, 1):", [(d[t, 22:26, 0] - t0).tolist() for t in range(NT)])
print("tile 0..5 e2a done (blocks 22, 23 | 0, 1):", [(d[t, 22:26, 3] - t0).tolist() for t in range(NT)])
print("tile 0 boundary (rel. clk): e2b(23) done, detect GEMM done, detect epilogue done, barrier+finalise done | next group: mel chunks stored, input GEMM done, first gate_wake:",
      (d[NT + 1, 20, :4] - t0).tolist(), (d[NT + 1, 24 + 21, :2] - t0).tolist(), int(d[0, 24, 0] - t0))
