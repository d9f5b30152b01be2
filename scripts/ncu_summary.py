#!/usr/bin/env python3
"""Dev tool: turn an .ncu-rep (ncu --set full) into the markdown summary kept under profiles/.
usage: ncu_summary.py report.ncu-rep "title" "command" > profiles/xxx.md"""
import csv, subprocess, sys, io
rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct', 'sm__throughput.avg.pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct', 'sm__pipe_fp64_cycles_active.avg.pct',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.per_cycle_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg.per_second', 'lts__t_sector_hit_rate.pct', 'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'smsp__inst_executed_op_shfl', 'sm__inst_executed_pipe_uniform')
print('# %s\n' % title)
print('Command: `%s`\n(numbers under the profiler are evidence of SHARES and counters, not bench values)\n' % cmd)
print('| metric | unit | value |\n|---|---|---|')
for k in sorted(d):
    if any(w in k for w in want) and 'pcsamp' not in k and '.max' not in k and '.min' not in k and (not k.endswith('.sum.per_second')):
        print('| %s | %s | %s |' % (k, d[k][0], d[k][1]))
