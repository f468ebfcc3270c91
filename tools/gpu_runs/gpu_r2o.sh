#!/bin/bash
set -x
for mode in default direct cap2; do
  case $mode in
    default) env="" ;;
    direct) env="CMSISDSP_CUDA_KERNEL=direct" ;;
    cap2) env="CMSISDSP_CUDA_PIPE_MAXOCC=2" ;;
  esac
  env $env python bench.py --steps 100 --no-e2e --no-secondary --no-cpu > gpurun_out/r2o_bench_$mode.json 2> gpurun_out/r2o_bench_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2o_bench_$mode.json").read().strip().splitlines()[-1])
r=d["roofline"]; s=r["sustained"]
print("$mode", "burst ms", round(d["ms_per_step"],4), "frac", round(r["frac"],3), "| sustained ms", round(s["ms_per_step_settled"],4), "frac", round(s["frac"],3), "sm_mhz", s["clocks"]["sm_mhz"], "power", s["clocks"]["power_w_max"], s["clocks"]["reasons"], "copy", round(s["sustained_copy_gbs"]))
PY
done 2>&1 | grep -v "^+" | tee gpurun_out/r2o_sustained_modes.txt
