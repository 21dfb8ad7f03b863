#!/bin/bash
# One box, N GPUs: the headline bench (BASELINE.json configs[1]) and the genome-scale workload (configs[2], >= 60 s steady state)
# at N ranks.  Usage (under gpurun --gpus N):  bash tools/scale_run.sh N [genome_steps]
set -u
N=${1:-1}
GSTEPS=${2:-220}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  RUN="python"
else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
fi
$RUN bench.py --gpus $N --steps 20 --warmup 3 --cpu-sample 0 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
$RUN bench.py --gpus $N --workload genome --steps $GSTEPS --warmup 3 --cpu-sample 0 > gpurun_out/r02_config3_n$N.json 2> gpurun_out/r02_config3_n$N.err
for f in gpurun_out/r02_scale_n$N.json gpurun_out/r02_config3_n$N.json; do
  python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value", round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "steps", d["steps"],
          "e2e", round(d["e2e"]["value"], 1), "clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
tail -2 gpurun_out/r02_config3_n$N.err
