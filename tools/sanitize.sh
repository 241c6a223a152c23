#!/bin/bash
# compute-sanitizer memcheck + racecheck (+ synccheck) over the GPU tests that exercise the kernels with hand-rolled
# synchronisation: per-warp mbarrier rings / bulk copies (gg_splat_tma.cu), the fused loss kernel, programmatic
# dependent launch chains, the peer-memory exchange (gg_p2p.cu, one-rank group).  Development tool; logs -> gpurun_out/.
#   gpurun --timeout 3000 -- bash tools/sanitize.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GG_SWEEP_CASES=3
SEL='nhwc or fused or p2p_shared or concat or multi_view or programmatic or cycle or bit_packed or config2 or needles or guard_boundary_sharpness'
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool" 
  timeout 1400 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 0 \
      python -m pytest tests -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/r2_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" gpurun_out/r2_sanitizer_$tool.log | tail -4
done
