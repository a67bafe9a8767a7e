"""Print the key metrics of an .ncu-rep (dev tool): python tools/ncu_key.py <report>"""
import csv
import subprocess
import sys

rows = list(csv.reader(subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'lts__t_sectors_srcunit_tex_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'smsp__inst_executed_pipe_fp64_op_dmma', 'sm__inst_executed_pipe_lsu']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for h, u in zip(hdr, units):
        if any(h == w or h.startswith(w) for w in want):
            print(f'{h} [{u}] = {d[h]}')
    print('-- warp stall reasons (per issue) --')
    st = [(h, float(d[h])) for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and d[h]]
    if not st:
        st = [(h, float(d[h])) for h in hdr if 'issue_stalled' in h and h.endswith('.ratio') and d[h]]
    for h, v in sorted(st, key=lambda x: -x[1])[:12]:
        print(f'  {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):32s} {v:.3f}')
