#!/bin/bash
# Runs of record on the B200 box (under gpurun): GPU test suite, bench lines, ncu launch lists and full captures.
# Usage: gpurun -- 'bash tools/run_of_record.sh r01b'      (outputs under gpurun_out/<tag>_*)
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $out/${tag}_tests.log 2>&1
(timeout 300 python bench.py 2>&1 | tail -1) > $out/${tag}_bench_config2.json
(timeout 300 python bench.py --workload config3 --steps 200 2>&1 | tail -1) > $out/${tag}_bench_config3.json
(timeout 300 python bench.py --workload config4 --steps 1024 --warmup 64 2>&1 | tail -1) > $out/${tag}_bench_config4_1gpu.json
(timeout 300 python bench.py --workload flows --steps 50 --warmup 5 2>&1 | tail -1) > $out/${tag}_bench_flows.json
(timeout 300 python bench.py --workload config5 --steps 100 --warmup 5 2>&1 | tail -1) > $out/${tag}_bench_config5.json
(timeout 300 python bench.py --impl reference --steps 100 --warmup 5 2>&1 | tail -1) > $out/${tag}_bench_reference_arm.json
(timeout 300 python bench.py --impl reference --workload flows --steps 10 --warmup 2 2>&1 | tail -1) > $out/${tag}_bench_reference_arm_flows.json
# launch lists (per-launch times are cold-cache and serialised: shares, not absolutes)
for w in config2 config3 flows; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_${w}_launches.csv \
      python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --only-device-pass > /dev/null 2>&1
done
# full captures of the dominant kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcc_step_warp -s 30 -c 1 -o $out/${tag}_config2_step \
    python bench.py --workload config2 --steps 40 --warmup 3 --no-cpu-baseline --only-device-pass > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcc_step_warp -s 30 -c 1 -o $out/${tag}_config3_step \
    python bench.py --workload config3 --steps 40 --warmup 3 --no-cpu-baseline --only-device-pass > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pcc_flows_ingest -s 3 -c 1 -o $out/${tag}_flows_ingest \
    python bench.py --workload flows --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cat $out/${tag}_tests.log
for f in config2 config3 config4_1gpu flows config5 reference_arm reference_arm_flows; do cut -c1-260 $out/${tag}_bench_$f.json; echo; done
