#!/bin/bash
# Config 5 evidence (under gpurun): bench line, ncu launch list of the bench command, ncu --set full of one step kernel.
# usage: gpurun --timeout 900 -- 'bash tools/ncu_config5.sh r02'
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
(timeout 400 python bench.py --workload config5 2>$out/${tag}_bench_config5.err | tail -1) > $out/${tag}_bench_config5.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_config5_launches.csv \
    python bench.py --workload config5 --steps 20 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# one step kernel of the same bench command (L2 flushed before it, as in the timed region)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcc_mwarp_step_kernel -s 25 -c 1 \
    -f -o $out/${tag}_config5_step python bench.py --workload config5 --steps 30 --warmup 5 --no-cpu-baseline > /dev/null 2>&1
cut -c1-1500 $out/${tag}_bench_config5.json
