"""Summary of one `ncu --set full --import-source on` capture: pipe utilisation, stall reasons per issued instruction, dynamic
opcode mix per work unit and a per-address-bin profile (instructions, stall samples).  usage: ncu_regions.py REPORT UNITS [BIN]"""
import csv, collections, sys, subprocess
rep=sys.argv[1]; frames=float(sys.argv[2]); B=int(sys.argv[3]) if len(sys.argv)>3 else 80
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h,u,v=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg']
for a,b,c in zip(h,u,v):
    if a in want: print(a,b,c)
for a,b,c in zip(h,u,v):
    if 'smsp__average_warps_issue_stalled' in a and 'per_issue_active' in a and float(c)>0.05: print(a.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),c)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]
iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iSamp=hdr.index('# Samples')
iw=hdr.index('stall_wait'); iss=hdr.index('stall_short_sb'); im=hdr.index('stall_math'); il=hdr.index('stall_long_sb')
S=sum(int(r[iSamp] or 0) for r in data); T=sum(int(r[iE] or 0) for r in data)
print("total warp-instr per unit", T/frames)
tot=collections.Counter()
for r in data:
    src_=r[iS].strip(); op=(src_.split()[1] if src_.startswith('@') else src_.split()[0]).rstrip(';')
    tot[op]+=int(r[iE] or 0)
print(' '.join('%s:%.1f'%(o,c/frames) for o,c in tot.most_common(28)))
for b in range(0,len(data),B):
    chunk=data[b:b+B]
    inst=sum(int(r[iE] or 0) for r in chunk)/frames
    if inst<1: continue
    samp=sum(int(r[iSamp] or 0) for r in chunk)
    w=sum(int(r[iw] or 0) for r in chunk); ss=sum(int(r[iss] or 0) for r in chunk); mm=sum(int(r[im] or 0) for r in chunk); ll=sum(int(r[il] or 0) for r in chunk)
    ops=collections.Counter()
    for r in chunk:
        src_=r[iS].strip(); op=(src_.split()[1] if src_.startswith('@') else src_.split()[0]).split('.')[0]
        ops[op]+=int(r[iE] or 0)
    top=' '.join('%s:%d'%(o,c/frames) for o,c in ops.most_common(6))
    print("%4d inst %6.1f samp %5.1f%% wait %4.1f ssb %4.1f math %4.1f lsb %4.1f | %s"%(b,inst,100*samp/S,100*w/S,100*ss/S,100*mm/S,100*ll/S,top))
