#!/bin/bash
# A/B bench helper (run through gpurun): each argument is "label:ENV=VAL,ENV=VAL" -> gpurun_out/ab_<label>.json + one summary line
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  env $(echo "$envs" | tr ',' ' ') python bench.py --steps ${AB_STEPS:-10} --warmup 3 --cpu-sample 0 ${AB_ARGS} > gpurun_out/ab_${label}.json 2> gpurun_out/ab_${label}.err
  python - "$label" <<PY
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    st=d.get("stage_ms_per_step",{})
    print(n, round(d["value"],1), round(d["ms_per_step"],2), {k:v for k,v in st.items() if v>0.5}, d["clocks"]["sm_mhz"], d.get("parity",{}).get("max_abs_err_scored"))
except Exception as e: print(n,"ERR",e, open(f"gpurun_out/ab_{n}.err").read()[-400:])
PY
done
