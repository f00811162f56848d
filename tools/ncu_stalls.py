"""Per-kernel stall summary from an .ncu-rep (source page):  python tools/ncu_stalls.py rep [kernel-substr] [topN] [occurrence]
(occurrence: which launch of that kernel in the report, default 0)"""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25; occ = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = re.split(r'(?m)^"Kernel Name",', out)
seen = collections.Counter()
for blk in blocks[1:]:
    lines = blk.split("\n")
    name = lines[0].strip().strip('",')
    if want not in name:
        continue
    seen[name] += 1
    if seen[name] - 1 != occ:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    h = rows[0]; idx = {k: i for i, k in enumerate(h)}; data = [r for r in rows[1:] if len(r) == len(h)]
    def f(r, k):
        try: return float(r[idx[k]])
        except Exception: return 0.0
    tot = sum(f(r, '# Samples') for r in data) or 1
    stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    print("==== %s   samples %d" % (name[:80], tot))
    agg = {k: sum(f(r, k) for r in data) for k in stalls}
    print("  " + "  ".join("%s %.2f" % (k[6:], v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:topn]:
        top = sorted(stalls, key=lambda k: -f(r, k))[:2]
        print('%7d %5.2f%% %-64s %s' % (f(r, '# Samples'), 100 * f(r, '# Samples') / tot, r[idx['Source']][:64], [(t[6:], int(f(r, t))) for t in top]))
    ops = collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']])
        if m: ops[m.group(2).split('.')[0]] += f(r, 'Instructions Executed')
    t = sum(ops.values()) or 1
    print("  instr mix: " + "  ".join("%s %.1f%%" % (k, 100 * v / t) for k, v in ops.most_common(14)), " total warp-instr %.3g" % t)
