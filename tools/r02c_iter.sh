#!/bin/bash
# r02c iteration: parity subset, bench line, per-part times, pair-schedule cost sweep, pair-kernel trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02c}
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${T}_tests.txt
for s in i8x3; do
  timeout 120 python bench.py --splits $s --steps 300 --warmup 10 --no-cpu-baseline --no-secondary > gpurun_out/${T}_b_$s.json 2> gpurun_out/${T}_b_$s.err || tail -3 gpurun_out/${T}_b_$s.err
  python - "$s" "$T" <<'PY'
import json, sys
s, T = sys.argv[1:3]
try:
    l = json.load(open(f"gpurun_out/{T}_b_{s}.json"))
    print(s, "us/round", round(l["ms_per_step"] * 1e3, 2), "frac", round(l["roofline"]["frac"], 3), "latency us", round(l["ms_per_round_latency"] * 1e3, 1),
          "e2e ms", round(l["e2e"]["ms_per_step"], 3), "clk", l["clocks"]["sm_mhz"], "parity", l["config"]["parity"]["max_abs_err_over_max_abs_ref"])
except Exception as e:
    print(s, "bench failed:", e)
PY
done
for c in "13,21,10" "8,12,6" "18,28,14" "4,4,0"; do
  echo "== H2_PAIR_COSTS=$c"
  H2_PAIR_COSTS=$c timeout 90 python tools/time_parts.py i8x3 2>&1 | grep -E "flush" | tee -a gpurun_out/${T}_costs.txt
  H2_PAIR_COSTS=$c timeout 90 python tools/pipeline_parts.py 2>&1 | grep lanes | tee -a gpurun_out/${T}_costs.txt
done
H2_BM_TRACE=1 python -m h2gcn_b200.build --force > /dev/null 2>&1
timeout 120 python tools/dbg_pair.py i8x3 2>&1 | head -24 | tee gpurun_out/${T}_trace_pair_i8x3.txt
python -m h2gcn_b200.build --force > /dev/null 2>&1
