#!/bin/bash
# Round-end evidence suite on one B200 (run through gpurun): bench (both arms), the other workloads, ncu launch list and
# captures, compute-sanitizer.  Everything lands in gpurun_out/; tools/summarize_profiles.py <tag> and tools/collect_results.py
# turn it into profiles/.     bash tools/final_suite.sh r02
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 180 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; head -c 300 gpurun_out/${TAG}_bench_default.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; head -c 200 gpurun_out/${TAG}_bench_ref.json; echo
python bench.py --steps 20 --warmup 3 --cpu-sample 0 > gpurun_out/${TAG}_scale_n1.json 2> /dev/null
[ -n "$SUITE_CONFIG3" ] && python bench.py --workload genome --steps 220 --cpu-sample 0 > gpurun_out/${TAG}_config3_n1.json 2> /dev/null
python bench.py --workload mutagenesis --cpu-sample 0 > gpurun_out/${TAG}_final_mutagenesis.json 2> /dev/null
python bench.py --workload long --model l32 --cpu-sample 0 > gpurun_out/${TAG}_final_long_l32.json 2> /dev/null
python bench.py --workload long --model cad2-small --cpu-sample 0 > gpurun_out/${TAG}_final_long_cad2small.json 2> /dev/null
python bench.py --workload long --model cad2-large --batch 8 --cpu-sample 0 > gpurun_out/${TAG}_final_long_cad2large.json 2> /dev/null
python tools/readme_benchmark.py --out gpurun_out/${TAG}_readme_benchmark.json > gpurun_out/${TAG}_readme_benchmark.log 2>&1; tail -4 gpurun_out/${TAG}_readme_benchmark.log
python tools/cli_vcf_benchmark.py --out gpurun_out/${TAG}_cli_vcf_benchmark.json > gpurun_out/${TAG}_cli_vcf_benchmark.log 2>&1; tail -1 gpurun_out/${TAG}_cli_vcf_benchmark.log
python tools/small_batch_probe.py > gpurun_out/${TAG}_small_batch.log 2>&1; tail -3 gpurun_out/${TAG}_small_batch.log
B="python bench.py --steps 1 --warmup 1 --cpu-sample 0 --no-clocks"
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:biscan -s 3 -c 1 -o gpurun_out/scan_${TAG} -f $B > gpurun_out/ncu_scan.log 2>&1
ncu --set full --clock-control none -k regex:gemm_bf16 -s 8 -c 6 -o gpurun_out/gemm_${TAG} -f $B > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none -k "regex:conv_silu|add_rmsnorm" -c 2 -o gpurun_out/elem_${TAG} -f $B > gpurun_out/ncu_elem.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:ssd_chunk_tc_kernel|gated_norm_sum" -s 4 -c 2 -o gpurun_out/ssd_${TAG} -f python bench.py --workload long --model cad2-small --batch 8 --steps 1 --warmup 1 --cpu-sample 0 --no-clocks > gpurun_out/ncu_ssd.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_ops_gpu.py tests/test_mamba2_gpu.py tests/test_model_gpu.py -m gpu -q -k "(biscan and not 8192) or conv or rmsnorm or (ssd_scan and not 8192 and not 640) or gated_norm or fp32_bc_rows" > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/${TAG}_sanitizer_memcheck.log | tail -3
compute-sanitizer --tool racecheck python -m pytest tests/test_ops_gpu.py tests/test_mamba2_gpu.py tests/test_model_gpu.py -m gpu -q -k "(biscan and 64-128) or (in_kernel_dt and 3-64-128) or (tcgen05 and 2-100-2) or (tcgen05 and 1-256-2) or (time_parallel and 1-512-128) or (fp32_bc_rows and 2-37-5)" > gpurun_out/${TAG}_sanitizer_racecheck.log 2>&1; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitizer_racecheck.log | tail -3
ls gpurun_out | grep ${TAG} | tr "\n" " "
