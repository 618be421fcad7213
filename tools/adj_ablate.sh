#!/bin/bash
# Timing experiments on the adjacency GEMM (ablated runs give wrong results by construction).  usage: tools/adj_ablate.sh "<env>" ...
mkdir -p gpurun_out
out=gpurun_out/adj_ablate.log
: > $out
for v in "$@"; do
  echo "=== variant: [$v]" >> $out
  env $v MDF_GEMM_TRACE=1 python tools/config_run.py --config 4 --n 16384 --check 0 --reps 1 2>&1 | grep -E "graphconv_adj|adj gemm trace" | tail -4 >> $out
done
cat $out
