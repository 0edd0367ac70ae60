#!/bin/bash
# compute-sanitizer over the one-link-per-warp multi-sender kernel (pcc_mwarp_step_kernel<S>, S = 2 and 3: the goldens).
# Usage: gpurun --timeout 1200 -- 'bash tools/sanitizer_multi.sh'   -> gpurun_out/r02_sanitizer_multi_*.log
out=gpurun_out
mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  (timeout 400 $S --tool $tool python -m pytest tests/test_gpu_multi.py -q -x -k "warp and golden" 2>&1 | tail -6) > $out/r02_sanitizer_multi_$tool.log
done
for f in $out/r02_sanitizer_multi_*.log; do tail -n 3 $f; done
