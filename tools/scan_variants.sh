#!/bin/bash
# Times the scan kernel of every library variant under variants/ (and the default build) on one GPU, same box, back to back.
mkdir -p gpurun_out
out=gpurun_out/scan_variants.txt
: > $out
for rep in 1 2; do
  echo "default $(python tools/bench_ops.py --ops scan 2>&1 | grep -o '"scan_ms": [0-9.]*')" >> $out
  for lib in variants/libpcad_*.so; do
    echo "$lib $(PCAD_LIB=$PWD/$lib python tools/bench_ops.py --ops scan 2>&1 | grep -o '"scan_ms": [0-9.]*')" >> $out
  done
done
cat $out
