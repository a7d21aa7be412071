"""Summarise an ncu report (raw page) + per-instruction source page into text kept under profiles/."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size']
for r in rows[2:]:
    print('=====', r[h.index('Kernel Name')][:60])
    for w in want:
        if w in h:
            print(f"  {w:70s} {r[h.index(w)]:>16s} {rows[1][h.index(w)]}")
    items = []
    for i, name in enumerate(h):
        if 'smsp__average_warps_issue_stalled' in name and name.endswith('_per_issue_active.ratio'):
            try:
                items.append((float(r[i]), name.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print('  warp stall reasons (warps per issue-active cycle):',
          ', '.join(f"{n}={v:.2f}" for v, n in sorted(items, reverse=True)[:8]))
