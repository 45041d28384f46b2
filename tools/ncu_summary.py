"""Print a compact table of the metrics we track from an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active']
extra = sys.argv[2:] 
names = [d[idx['Kernel Name']].replace('void ', '').split('(')[0] for d in data]
print("| metric | unit | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for k in KEYS + extra:
    if k in idx:
        print(f"| {k} | {units[idx[k]]} | " + " | ".join(d[idx[k]] for d in data) + " |")
for h in hdr:
    if 'issue_stalled' in h and h.endswith('_per_warp_active.pct'):
        vals = [d[idx[h]] for d in data]
        try:
            if max(float(v) for v in vals) >= 4.0:
                print(f"| {h} | % | " + " | ".join(vals) + " |")
        except ValueError:
            pass
