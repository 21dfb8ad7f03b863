#!/bin/bash
# One 8-GPU box, same lease: the headline bench at N = 1 and N = 8 back to back (so the ratio is not across boxes with different
# power-capped clocks) and config 3 at N = 1.      under gpurun --gpus 8:  bash tools/scale_ladder.sh
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
python bench.py --gpus 1 --steps 20 --warmup 3 --cpu-sample 0 > gpurun_out/r02_ladder_n1.json 2> gpurun_out/r02_ladder_n1.err
$T --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 3 --cpu-sample 0 > gpurun_out/r02_ladder_n8.json 2> gpurun_out/r02_ladder_n8.err
$T --nproc-per-node 4 bench.py --gpus 4 --steps 20 --warmup 3 --cpu-sample 0 > gpurun_out/r02_ladder_n4.json 2> gpurun_out/r02_ladder_n4.err
python bench.py --gpus 1 --workload genome --steps 70 --warmup 3 --cpu-sample 0 > gpurun_out/r02_ladder_config3_n1.json 2> /dev/null
for f in gpurun_out/r02_ladder_*.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
