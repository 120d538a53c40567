"""Summarise an .ncu-rep: key raw metrics + SASS instruction mix of the first kernel."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
 'lts__t_bytes.sum','l1tex__t_bytes.sum','launch__grid_size','launch__block_size','sm__cycles_elapsed.avg','launch__occupancy_limit_registers',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
for r in rows[2:3]:
    print('kernel:', r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr: print(f'  {w:80s} {r[hdr.index(w)]} {units[hdr.index(w)]}')
    for h in hdr:
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') or ('warp_issue_stalled' in h and h.endswith('.pct')):
            pass
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr=None; data=[]
for r in rows:
    if r and r[0]=='Address':
        if hdr is not None: break
        hdr=r; continue
    if hdr is not None and len(r)==len(hdr): data.append(r)
ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
tot=sum(int(r[ie]) for r in data); nw = max(int(r[ie]) for r in data)
print('sass lines',len(data),'warp-instr',tot,'per warp',tot/nw)
c=collections.Counter(); s=collections.Counter()
for r in data:
    t=r[ia].split(); op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]; c[op]+=int(r[ie]); s[op]+=int(r[isamp])
print(' '.join(f'{op}:{n/nw:.0f}' for op,n in c.most_common(45)))
open(rep+'.sass.txt','w').write('\n'.join(f"{k:5d} {int(r[ie]):7d} {int(r[isamp]):4d}  {r[ia].strip()}" for k,r in enumerate(data)))
# stall columns
st=[h for h in hdr if h.startswith('stall_') or 'Stall' in h]
