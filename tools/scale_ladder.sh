#!/bin/bash
# One 8-GPU box: headline bench and config 3 at N = 1, 2, 4, 8 back to back (same box, so the ratios are clean).
#   under gpurun --gpus 8:  bash tools/scale_ladder.sh
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = "8" ]; then G=220; else G=70; fi      # config 3: >= 60 s steady state at the full box, ~20 s at the smaller counts
  bash tools/scale_run.sh $N $G 2>&1 | grep -E "^gpurun_out|ERR"
done
