#!/usr/bin/env bash
# TMA ragged-copy microbenchmark (tools/tma_ragged_copy.cu): tensor-map boxes at arbitrary element offsets
set -u
mkdir -p gpurun_out
{
for m in 1 2 3; do echo "mode $m"; timeout 120 build/tma_ragged_copy 1366 768 4 $m; done
echo "sanitizer"; timeout 200 compute-sanitizer --print-limit 3 build/tma_ragged_copy 1366 64 4 0 2>&1 | head -30
} 2>&1 | tee gpurun_out/tma_micro.log
