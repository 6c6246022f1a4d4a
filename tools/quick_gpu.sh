#!/bin/bash
# Quick on-box check used while tuning the tensor-core path: parity subset, bench lines, per-part timings.
#   tools/quick_gpu.sh [splits ...]   (default: i8x2 i8x3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core or int8" 2>&1 | tail -2
for s in ${@:-i8x2 i8x3}; do
  timeout 90 python bench.py --splits $s --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/b_$s.json 2> gpurun_out/b_$s.err || tail -3 gpurun_out/b_$s.err
  python - "$s" <<'PY'
import json, sys
s = sys.argv[1]
try:
    l = json.load(open(f"gpurun_out/b_{s}.json"))
    print(s, "us/round", round(l["ms_per_step"] * 1e3, 2), "frac", round(l["roofline"]["frac"], 3), "flush+events",
          round(l["config"]["ms_per_step_l2_flush_events"] * 1e3, 1), "e2e ms", round(l["e2e"]["ms_per_step"], 3), "clk", l["clocks"]["sm_mhz"])
except Exception as e:
    print(s, "bench failed:", e)
PY
  timeout 60 python tools/time_parts.py $s 2>&1 | grep -E "tensor|csr only"
done
