"""Summarise an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u = r[0], r[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster',
        'smsp__inst_executed.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_red.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'launch__shared_mem_per_block_dynamic', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled']
for v in r[2:]:
    print('KERNEL', v[h.index('Kernel Name')], file=out)
    for i, name in enumerate(h):
        if any(name == k or (k.endswith('stalled') and name.startswith(k)) or name == k + ' ' for k in KEYS):
            if name.startswith('smsp__average_warps_issue_stalled') and float(v[i] or 0) < 0.3:
                continue
            print('   %-90s %-12s %s' % (name, u[i], v[i]), file=out)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh, data = rows[1], rows[2:]
isrc, ist, iex = hh.index('Source'), hh.index('Warp Stall Sampling (All Samples)'), hh.index('Instructions Executed')
tot = sum(int(x[ist]) for x in data) or 1
print('\nSASS stall samples: total %d over %d instructions; blocks of 40 instructions with >1%%:' % (tot, len(data)), file=out)
for b in range(0, len(data), 40):
    s = sum(int(x[ist]) for x in data[b:b + 40])
    if s > tot * 0.01:
        print('  [%4d..] %5.1f%%  %s' % (b, 100 * s / tot, data[b][isrc].strip()[:70]), file=out)
print('top instructions:', file=out)
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][ist]))[:30]):
    x = data[i]
    print('  %4d %-84s %6s %5.1f%% exec %s' % (i, x[isrc].strip()[:84], x[ist], 100 * int(x[ist]) / tot, x[iex]), file=out)
