#!/bin/bash
# Runs of record of round 2 on the B200 box (under gpurun): GPU test suite, smoke, bench lines, config 5 evidence, ncu
# launch list and full capture of the dominant kernel, drop-in timing.  Most important first: a call that runs out of
# box time still leaves the earlier outputs.  (clock64 phase profiles: tools/build_profile_lib.sh + phase_profile_*.py.)
# Usage: gpurun --timeout 1500 -- 'bash tools/run_of_record_r02.sh r02'      (outputs under gpurun_out/<tag>_*)
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $out/${tag}_tests.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > $out/${tag}_smoke.txt
(timeout 600 python bench.py 2>$out/${tag}_bench_config3.err | tail -1) > $out/${tag}_bench_config3.json
# config 5 (several senders per link): bench line, launch list, full capture of one step kernel
bash tools/ncu_config5.sh $tag > /dev/null 2>&1
(timeout 300 python bench.py --impl reference --steps 100 --warmup 5 2>/dev/null | tail -1) > $out/${tag}_bench_reference_arm.json
(timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1) > $out/${tag}_bench_config3_k20.json
(timeout 300 python bench.py --workload config2 --no-cpu-baseline 2>/dev/null | tail -1) > $out/${tag}_bench_config2.json
(timeout 300 python bench.py --workload config4 --steps 2048 2>/dev/null | tail -1) > $out/${tag}_bench_config4_1gpu.json
(timeout 120 python tools/time_dropin.py 10 2>&1 | tail -2) > $out/${tag}_dropin.txt
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_config3_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --only-device-pass > /dev/null 2>&1
# full capture of the dominant kernel, mid-episode
EXP_WARMUP=185 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcc_step_packed_kernel -s 195 -c 1 \
    -f -o $out/${tag}_config3_step python tools/exp_modes.py 65536 PCC_B200_MODE=packed > /dev/null 2>&1
cat $out/${tag}_tests.log $out/${tag}_dropin.txt $out/${tag}_smoke.txt
for f in config3 config3_k20 config2 config4_1gpu reference_arm config5; do cut -c1-400 $out/${tag}_bench_$f.json; echo; done
