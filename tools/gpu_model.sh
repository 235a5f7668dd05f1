#!/bin/bash
# model-level checks on the GPU box
timeout 600 python -m pytest tests/test_model_host.py -m gpu -x -q 2>&1 | tail -4
for wl in "xfmamba_t_infer --batch 64" "xfmamba_s_infer --batch 64" "xfmamba_s_infer --batch 64 --no-graph" "xfmamba_b_train --batch 32" "xfmamba_b_hires --batch 8 --no-graph"; do
  tag=$(echo $wl | tr ' -' '__')
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/model_${tag}.json 2> gpurun_out/model_${tag}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/model_${tag}.json"))
    print("${wl}: %.1f pairs/s  %.2f ms/step | e2e %.1f pairs/s | launches %d | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"]["mode"]))
except Exception as e:
    print("${wl}: FAILED", e); print(open("gpurun_out/model_${tag}.err").read()[-1500:])
PY
done
