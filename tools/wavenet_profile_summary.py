"""profiles/r2_wavenet_ncu_summary.txt + profiles/r2_traffic.json from one `ncu --set full` capture of the two WaveNet launches
(stream-level pass, window pass) of:  bench.py --workload wavenet --streams 512 --steps 1 --warmup 3 --no-extras --no-cpu-baseline
    python tools/wavenet_profile_summary.py gpurun_out/<capture>.ncu-rep"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
out = ["# wavenet_tc_kernel, ncu --set full --clock-control none, 512 streams x 10 s (208 896 hop-2 windows); id 0 = stream-level snapshot pass, "
       "id 1 = window pass (shared activations).  Round-2 final kernel: 4 windows on 6 tiles, 832 threads, skip sums in shared memory, snapshots as column planes.",
       "# command: ncu --set full --clock-control none --import-source on -k regex:wavenet_tc -s 6 -c 2 python bench.py --workload wavenet --streams 512 --steps 1 --warmup 3 --no-extras --no-cpu-baseline"]
for ri, r in enumerate(rows[2:]):
    out.append("---- %s id %d" % (r[h.index('Kernel Name')][:40], ri))
    for a, b, c in zip(h, u, r):
        if a in want:
            out.append("   %s %s %s" % (a, b, c))
    st = [(a.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(c)) for a, c in zip(h, r)
          if 'smsp__average_warps_issue_stalled' in a and 'per_issue_active' in a and 'not_issued' not in a]
    out.append("   stalls per issue: " + ", ".join("%s %.2f" % (a, c) for a, c in sorted(st) if c > 0.09))
open(os.path.join(ROOT, "profiles", "r2_wavenet_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
g = lambda r, k: float(r[h.index(k)])
assert u[h.index('dram__bytes_read.sum')] == 'Mbyte'
sp, wp = rows[2], rows[3]
rd, wr = g(wp, 'dram__bytes_read.sum'), g(wp, 'dram__bytes_write.sum')
j = {"wavenet encode+detect": {"streams": 512, "seconds": 10.0, "bytes_per_launch": int(round((rd + wr) * 1e6)), "scales_with": "streams",
     "note": "dram__bytes_read.sum + dram__bytes_write.sum of the WaveNet window pass (wavenet_tc_kernel, id 1 of profiles/r2_wavenet_ncu_summary.txt): "
             "%.1f MB read (per-frame snapshots of the shared activations, 5 levels x 192 B per frame = 490 MB per launch, + the 33 MB per-row input layer) "
             "+ %.1f MB written; the stream-level pass that writes the snapshots moves another %.0f MB in %.2f ms"
             % (rd, wr, g(sp, 'dram__bytes_read.sum') + g(sp, 'dram__bytes_write.sum'), g(sp, 'gpu__time_duration.sum'))}}
json.dump(j, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print("\n".join(out))
