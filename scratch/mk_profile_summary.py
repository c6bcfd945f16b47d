"""Condense ncu reports into the small text/CSV files committed under profiles/."""
import collections, csv, json, re, subprocess, sys
tag, fwd_rep, bwd_rep, launches = sys.argv[1:5]
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 64
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_shared_mem','smsp__warps_active.avg.per_cycle_active',
 'smsp__warps_eligible.avg.per_cycle_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct']
out = {'batch': batch}
for name, rep in (('fwd', fwd_rep), ('bwd', bwd_rep)):
    raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    r = list(csv.reader(raw.splitlines())); h, u, v = r[0], r[1], r[2]
    d = {'kernel': v[h.index('Kernel Name')]}
    for k in KEYS:
        if k in h: d[k] = f"{v[h.index(k)]} {u[h.index(k)]}".strip()
    st = {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''): float(v[h.index(k)])
          for k in h if 'issue_stalled' in k and k.endswith('ratio') and v[h.index(k)] not in ('', 'n/a')}
    d['stall_cycles_per_issued_instruction'] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:9])
    src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(src.splitlines())); hh = rows[1]; data = rows[2:]
    ia = hh.index('Instructions Executed'); isrc = hh.index('Source')
    ops = collections.Counter(); tot = 0
    for x in data:
        if not x[ia].isdigit(): continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', x[isrc]); op = m.group(2).split('.')[0] if m else '?'
        ops[op] += int(x[ia]); tot += int(x[ia])
    d['executed_warp_instructions_by_opcode_pct'] = {op: round(c / tot * 100, 1) for op, c in ops.most_common(18)}
    d['sass_has'] = {m: any(m in x[isrc] for x in data) for m in ('UBLKCP', 'SYNCS', 'FFMA2', 'FMUL2', 'MUFU.EX2', 'SHFL', 'REDUX')}
    out[name] = d
rows = list(csv.reader(open(launches)))
hdr = None; agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); k = re.sub(r'\(.*', '', d['Kernel Name'])[:70]
        agg[k][0] += 1; agg[k][1] += float(d['Metric Value'].replace(',', ''))
tot = sum(v[1] for v in agg.values())
out['launch_list'] = [{'kernel': k, 'launches': v[0], 'total_us': round(v[1] / 1e3, 1), 'share_pct': round(v[1] / tot * 100, 1),
                       'avg_us': round(v[1] / v[0] / 1e3, 1)} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]]
json.dump(out, open(f'profiles/{tag}_ncu_summary.json', 'w'), indent=1)
print(json.dumps(out, indent=1)[:3000])
