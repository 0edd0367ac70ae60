#!/bin/bash
# ncu --set full capture of ONE launch of a step kernel.  usage: tools/ncu_step.sh <kernel-regex> <out-name> <n_envs> <setting>
set -e
K="$1"; OUT="$2"; N="$3"; SETTING="$4"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name "regex:$K" --launch-skip 25 --launch-count 1 \
    -f -o gpurun_out/$OUT python tools/exp_modes.py $N "$SETTING" > gpurun_out/$OUT.log 2>&1
tail -3 gpurun_out/$OUT.log
