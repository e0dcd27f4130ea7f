#!/bin/bash
# K1 prefilter engine: launch-list breakdown + floor experiments (run under gpurun)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prefilter|rescore|exact_scan|gather|decode|prep" \
    --csv --log-file gpurun_out/tcp_launches.csv python bench.py --engine prefilter --profile --steps 1 --warmup 1 > /dev/null 2>&1
python -c "import sys; sys.path.insert(0,\"tools\"); import summarize_ncu as s; print(s.launches(\"gpurun_out/tcp_launches.csv\"))"
for cfg in "0 0" "4 0" "1 0" "2 0" "0 4" ; do
  set -- $cfg
  echo "EXP=$1 BH=$2: $(FGVC_TCP_EXP=$1 FGVC_TCP_BH=$2 python bench.py --engine prefilter --profile --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-200)"
done
