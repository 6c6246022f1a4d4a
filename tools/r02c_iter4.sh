#!/bin/bash
# FH = 32 at d = 128 (80 items of 157 half-cost units) under the current stream-K schedule: trace + pipelined rate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02g}
export H2_BM_FH32_UPTO=128
H2_BM_TRACE=1 python -m h2gcn_b200.build --force > /dev/null 2>&1
for c in "16,42,12" "26,26,0"; do
  echo "== FH32 trace H2_PAIR_COSTS=$c" | tee -a gpurun_out/${T}_traces.txt
  H2_TRACE_UNITS=40 H2_TRACE_DUMP=gpurun_out/${T}_ctas_${c//,/_}.csv H2_PAIR_COSTS=$c timeout 120 python tools/dbg_pair.py i8x3 2>&1 | head -64 | tee -a gpurun_out/${T}_traces.txt
done
python -m h2gcn_b200.build --force > /dev/null 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor or int8 or i8 or pair or bitmap or round" 2>&1 | tail -2 | tee gpurun_out/${T}_tests.txt
for c in "16,42,12" "26,26,0"; do
  echo "== FH32 H2_PAIR_COSTS=$c" | tee -a gpurun_out/${T}_costs.txt
  H2_PAIR_COSTS=$c timeout 90 python tools/pipeline_parts.py 2>&1 | grep "lanes 2" | tee -a gpurun_out/${T}_costs.txt
done
