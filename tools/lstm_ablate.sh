#!/bin/bash
# Timing experiments on the fused LSTM kernel (results of ablated runs are wrong by construction): one configs[4] chunk per variant.
# usage: tools/lstm_ablate.sh "<env assignments>" ...   (each argument = one variant, "" = baseline)
mkdir -p gpurun_out
out=gpurun_out/lstm_ablate.log
: > $out
for v in "$@"; do
  echo "=== variant: [$v]" >> $out
  env $v python tools/config_run.py --config 4 --n 16384 --check 0 --reps 2 2>&1 | grep -E "tc run 1|lstm|trace" >> $out
done
cat $out
