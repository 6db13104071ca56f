"""Summarise an ncu report (raw page) / launch list into a small text table for profiles/."""
import csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    h = rows[hdr]
    ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    out = []
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            out.append((r[ki], float(r[vi].replace(',', ''))))
    return out

def raw(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
            'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__cycles_active.avg',
            'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
            ]
    idx = {w: h.index(w) for w in want if w in h}
    res = []
    for r in rows[2:]:
        res.append({w: r[i] for w, i in idx.items()} | {'_units': {w: units[i] for w, i in idx.items()}})
    return res

if __name__ == '__main__':
    if sys.argv[1].endswith('.csv'):
        for k, v in launches(sys.argv[1]):
            if 'tdx::' in k:
                print(f"{v/1e3:10.1f} us  {k[:110]}")
    else:
        for r in raw(sys.argv[1]):
            u = r.pop('_units')
            print('---', r['Kernel Name'][:100])
            for k, v in r.items():
                if k != 'Kernel Name':
                    print(f"   {k:95s} {v:>16s} {u[k]}")
