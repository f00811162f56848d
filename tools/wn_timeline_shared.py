"""(needs a library built with -DWWB_TIMELINE [-DWWB_WN_DBG2]: tools/build_variant.sh tl wavenet_tc.cu -DWWB_TIMELINE, run with WWB200_LIB=build/libwwb200_tl.so) Timeline of the WaveNet kernel in shared-activation mode (sliding windows): period per block, per tile events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wakeword_detection_b200 import _cabi, weights as W
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = _cabi.Engine(W.load_model_dir(os.path.join(ROOT, "weights", "Wavenet"), "Wavenet"), 0, sys.argv[1] if len(sys.argv) > 1 else "tc")
mel = torch.rand((16, 998, 40), device=eng.device) * 5
eng.posteriors(mel, hop=2); torch.cuda.synchronize()
NT = int(os.environ.get('WN_NT', '6'))
NR = NT + 3
dbg = torch.zeros(NR * 48 * 4 + 64 + 24 * 16, dtype=torch.int64, device=eng.device)
eng.lib.wwb_debug_buffer(eng.ctx, dbg.data_ptr())
eng.posteriors(mel, hop=2); torch.cuda.synchronize()
d = dbg.cpu().numpy()[:NR * 48 * 4].reshape(NR, 48, 4)
t0 = d[NT, 0, 0]
np.set_printoptions(linewidth=220)
print("period per block (gate issue of tile 0), two consecutive groups:", np.diff(d[NT, :, 0]))
for k in (0, 3, 6, 8, 11, 14, 18, 20, 23):
    print("block", k, "issuer G0,G4,R0,R4:", d[NT, k] - t0)
    for tile in range(NT):
        e = d[tile, k]
        if e[0] == 0:
            continue
        print("   tile", tile, "gate_done, e1_done, rs_done, e2a_done:", e - t0, " e1=%d rs_wait=%d e2a=%d" % (e[1] - e[0], e[2] - e[1], e[3] - e[2]))
print("tile 0 boundary: e2b(23) done, detect GEMM done, detect epilogue done, barrier done:", (d[NT + 1, 20, :4] - t0).tolist(), " next group first gate issue:", int(d[NT, 24, 0] - t0))

f = dbg.cpu().numpy()[NR * 48 * 4 + 64:].reshape(24, 16)
names = ["gate_wake", "ld1", "math_done", "g_st_done", "arrive_g", "rs_wake", "x_upd", "sts", "u_st_done", "arrive_u",
         "G:wait", "G:woke", "G:issued", "R:woke", "R:issued"]
for k in (1, 2, 3, 4):
    order = [12, 0, 1, 2, 3, 4, 13, 14, 5, 6, 7, 8, 9]
    base = f[k, 12]
    print("block", k, "chain of tile 0 (clk after its gate issue):", ", ".join("%s %d" % (names[i], f[k, i] - base) for i in order),
          "| next: G:woke %d G:issued %d" % (f[k + 1, 11] - base, f[k + 1, 12] - base))
