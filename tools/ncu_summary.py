"""Summarise an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv > raw.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('---', r[idx['Kernel Name']], r[idx.get('Grid Size', 0)] if 'Grid Size' in idx else '')
    for w in want:
        if w in idx and r[idx[w]] not in ('', 'n/a'):
            print(f'  {w}: {r[idx[w]]} {units[idx[w]]}')
    st = {}
    for h in hdr:
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try: st[h] = float(r[idx[h]].replace(',', ''))
            except ValueError: pass
    tot = sum(st.values()) or 1
    print('  stalls (share of warp-cycles per issue):', ', '.join(
        f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}:{v / tot:.2f}"
        for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
