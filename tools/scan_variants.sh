#!/bin/bash
# Times (and parity-tests) scan-kernel build variants under variants/ x warp-flavour mixes on one GPU.
mkdir -p gpurun_out
out=gpurun_out/scan_variants.txt
: > $out
run() {  # lib mix
  echo "== $1 mix=$2" >> $out
  PCAD_SCAN_MIX=$2 PCAD_LIB=$PWD/variants/$1 python tools/bench_ops.py --ops scan --batch 256 2>&1 | grep -o '"scan_ms": [0-9.]*' >> $out
}
tst() {
  PCAD_SCAN_MIX=$2 PCAD_LIB=$PWD/variants/$1 python -m pytest tests/test_ops_gpu.py -q -m gpu -k biscan -x 2>&1 | tail -1 >> $out
}
for m in 0 0x10012 0x10112 0x01013 0x11113 0x21114 0x01014 0x43138 0x21125; do run libpcad_m_p0_u4_d4_h8.so $m; done
tst libpcad_m_p0_u4_d4_h8.so 0x11113
for m in 0x11113 0x21114 0x21125; do run libpcad_m_p0_u4_d3_h8.so $m; done
tst libpcad_m_p0_u4_d3_h8.so 0x11113
for m in 0x10012 0x10112 0x11113; do run libpcad_m_p0_u4_d4_h6.so $m; done
for m in 0x11113 0x21114; do run libpcad_m_p1_u4_d4_h8.so $m; done
cat $out
