#!/bin/bash
# Round-end measurement suite on one B200 (run through gpurun): tests, bench (both arms), ncu launch list and captures,
# compute-sanitizer.  Everything lands in gpurun_out/; tools/summarize_profiles.py turns it into profiles/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -2 gpurun_out/pytest_gpu_final.log
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; head -c 300 gpurun_out/bench_r1_final.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; head -c 300 gpurun_out/bench_r1_ref.json; echo
B="python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-clocks"
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:biscan -s 3 -c 1 -o gpurun_out/scan_r1 -f $B > gpurun_out/ncu_scan.log 2>&1
ncu --set full --clock-control none -k regex:gemm_bf16 -s 8 -c 6 -o gpurun_out/gemm_r1 -f $B > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none -k "regex:conv_silu|add_rmsnorm" -c 2 -o gpurun_out/elem_r1 -f $B > gpurun_out/ncu_elem.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_ops_gpu.py -m gpu -q -k "biscan or conv or rmsnorm or softplus" > gpurun_out/sanitizer_memcheck.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log | tail -3
compute-sanitizer --tool racecheck python -m pytest tests/test_ops_gpu.py -m gpu -q -k "biscan and 64-128" > gpurun_out/sanitizer_racecheck.log 2>&1; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck.log | tail -3
ls -la gpurun_out | tail -20
