#!/bin/bash
# one optimisation iteration on the GPU box: parity tests (fused subset), bench line summary, optional ncu capture
# usage: tools/gpu_iter.sh <tag> [ncu]
TAG=${1:-iter}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q  2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("${TAG}: pairs/s %.0f  step %.3f ms | bwd %.3f ms (%.1f%% of peak) | fwd %.3f ms (%.1f%% of peak) | e2e %.0f | clocks %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["avg_ms"], 100 * d["roofline"]["frac"], d["roofline_fwd"]["avg_ms"],
    100 * d["roofline_fwd"]["frac"], d["e2e"]["value"], d["clocks"]))
PY
tail -2 gpurun_out/bench_${TAG}.err
if [ "$2" == "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:ss2d -s 6 -c 2 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${TAG}.log 2>&1
  tail -1 gpurun_out/ncu_${TAG}.log
fi
