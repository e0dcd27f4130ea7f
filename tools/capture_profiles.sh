#!/bin/bash
# round-end evidence (run under gpurun): launch list + one full capture of the dominant kernel + bench line
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:affinity_topk_tc16 -s 1 -c 1 -f -o gpurun_out/k1_full \
    python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/launches.csv gpurun_out/k1_full.ncu-rep gpurun_out/summary.md "$1"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 2500 gpurun_out/bench_n1.json
