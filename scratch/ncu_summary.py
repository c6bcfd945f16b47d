import csv,sys,subprocess,collections,re
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(raw.splitlines())); h=r[0]; v=r[2] if len(r)>2 else r[1]
def g(k): return v[h.index(k)] if k in h else None
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active']
for k in keys: print(f'{k:80s} {g(k)}')
for k in h:
    if 'warp_issue_stalled' in k and 'per_warp_active.pct' in k:
        val=float(g(k)) if g(k) else 0
        if val>2: print(f'  {k:78s} {val:.1f}')
for k in h:
    if 'pipe' in k and 'pct_of_peak_sustained_active' in k and 'inst_executed' in k:
        try:
            val=float(g(k))
            if val>5: print(f'  {k:78s} {val:.1f}')
        except: pass
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hh=rows[1]; data=rows[2:]
ia=hh.index('Instructions Executed'); isrc=hh.index('Source'); isamp=hh.index('# Samples')
tot=sum(int(x[ia]) for x in data if x[ia].isdigit())
ops=collections.Counter()
for x in data:
    if not x[ia].isdigit(): continue
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', x[isrc]); op=m.group(2).split('.')[0] if m else '?'
    ops[op]+=int(x[ia])
print('total warp-inst',tot)
print(' '.join(f'{op}:{c/tot*100:.1f}%' for op,c in ops.most_common(28)))
if len(sys.argv)>2:
    open(sys.argv[2],'w').write(src)
