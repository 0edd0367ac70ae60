#!/bin/bash
# compute-sanitizer over the kernels new in round 2 (packed step kernel in all three roles, MT19937 solo warp, rollout
# sequence).  Usage: gpurun --timeout 1800 -- 'bash tools/sanitizer_r02.sh'   -> gpurun_out/r02_sanitizer_*.log
out=gpurun_out
mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
(timeout 600 $S --tool memcheck python -m pytest tests/test_gpu_packed.py -q -x -k "ragged or (mixed and bufferbloat) or (quad and heavyloss)" 2>&1 | tail -6) > $out/r02_sanitizer_memcheck_packed.log
(timeout 600 $S --tool memcheck python -m pytest tests/test_gpu_dropin.py tests/test_gpu_rollout.py -q -x -k "kat or value_head" 2>&1 | tail -6) > $out/r02_sanitizer_memcheck_mt_rollout.log
(timeout 600 $S --tool synccheck python -m pytest tests/test_gpu_packed.py -q -x -k "ragged" 2>&1 | tail -6) > $out/r02_sanitizer_synccheck_packed.log
(timeout 900 $S --tool racecheck python -m pytest tests/test_gpu_packed.py -q -x -k "mixed and tinyqueue" 2>&1 | tail -8) > $out/r02_sanitizer_racecheck_packed.log
(timeout 900 $S --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "bufferbloat" 2>&1 | tail -8) > $out/r02_sanitizer_racecheck_pair.log
tail -4 $out/r02_sanitizer_*.log
