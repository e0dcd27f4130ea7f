#!/bin/bash
# ncu --set full of the non-K1 kernels of one bench step (run under gpurun) -> gpurun_out/small_kernels.md
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:"prep_features|decode_argmax_jobs|decode_minmax_jobs|gather_chain|gather_weights" \
    -c 6 -f -o gpurun_out/small_full python bench.py --profile --steps 1 --warmup 0 > /dev/null 2>&1
ncu -i gpurun_out/small_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin)); h, u = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__registers_per_thread']
print('| kernel | ' + ' | '.join(k.split('.')[0] for k in keys) + ' |')
print('|---|' + '---:|' * len(keys))
for r in rows[2:]:
    name = r[h.index('Kernel Name')].split('(')[0][-40:]
    print('| \`' + name + '\` | ' + ' | '.join((r[h.index(k)] + ' ' + u[h.index(k)]) if k in h else '-' for k in keys) + ' |')
" | tee gpurun_out/small_kernels.md
