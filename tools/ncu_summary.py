"""Summarise an .ncu-rep (ncu -i ... --page raw --csv) into a small CSV of the metrics that matter here."""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
        'l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum', 'sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum']

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + [f'launch{i}' for i in range(len(rows) - 2)])
    for k in KEEP:
        if k in idx:
            w.writerow([k, units[idx[k]]] + [r[idx[k]] for r in rows[2:]])
for line in open(out):
    print(line.rstrip()[:260])
# DRAM traffic per launch, keyed by kernel name (bench.py reads roofline.traffic from this file)
if len(sys.argv) > 3 and 'dram__bytes_read.sum' in idx:
    import json

    def to_bytes(v, unit):
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)
        return float(v.replace(',', '')) * mult
    tr = {}
    for r in rows[2:]:
        name = r[idx['Kernel Name']]
        b = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        tr.setdefault(name, []).append({'grid': r[idx['Grid Size']], 'dram_bytes': b, 'time_ms': r[idx['gpu__time_duration.sum']]})
    json.dump(tr, open(sys.argv[3], 'w'), indent=1)
