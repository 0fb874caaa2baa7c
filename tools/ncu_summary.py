"""Summarise an .ncu-rep (read here, no GPU): per captured launch the metrics the roofline needs."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'launch__shared_mem_per_block_dynamic', 'smsp__cycles_active.avg']
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print('| # | kernel | grid | block | ' + ' | '.join(w.split('.')[0].replace('__', ':') + ('.' + w.split('.')[-1] if 'pct' in w else '') for w in WANT) + ' |')
for n, r in enumerate(rows[2:]):
    name = r[hdr.index('Kernel Name')]
    name = name.split('(')[0].split('::')[-1][:34]
    vals = []
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            vals.append(f'{r[i]} {units[i]}'.strip())
        else:
            vals.append('-')
    print(f'| {n} | {name} | {r[hdr.index("Grid Size")]} | {r[hdr.index("Block Size")]} | ' + ' | '.join(vals) + ' |')
