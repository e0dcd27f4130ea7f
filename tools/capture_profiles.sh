#!/bin/bash
# round evidence (run under gpurun, ONE GPU): launch list of the bench command + one full capture of the dominant kernel
# (single-CTA tiles: ncu 2025.x segfaults while replaying the cta_group::2 cluster kernel) + tail kernels + bench line.
# Every ncu call is bounded by `timeout`: a hung replay must not eat the GPU budget.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:affinity_topk_tc16 -s 1 -c 1 -f -o gpurun_out/k1_full \
    python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"prep_features|decode_argmax_jobs|gather_chain|gather_weights|decode_minmax_jobs" \
    -s 5 -c 5 -f -o gpurun_out/tail_full python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1
python tools/summarize_ncu.py gpurun_out/launches.csv gpurun_out/k1_full.ncu-rep gpurun_out/summary.md "$1"
ls -la gpurun_out/*.ncu-rep
