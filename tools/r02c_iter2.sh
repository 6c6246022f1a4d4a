#!/bin/bash
# r02c iteration 2: parity subset for the tensor path, per-part times under a few schedule costs, pair-kernel trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02d}
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor or int8 or i8 or pair or bitmap or round" 2>&1 | tail -3 | tee gpurun_out/${T}_tests.txt
for c in ${COSTS:-"13,21,10" "8,12,6" "10,10,4" "6,8,2"}; do
  echo "== H2_PAIR_COSTS=$c" | tee -a gpurun_out/${T}_costs.txt
  H2_PAIR_COSTS=$c timeout 90 python tools/time_parts.py i8x3 2>&1 | grep -E "flush" | tee -a gpurun_out/${T}_costs.txt
  H2_PAIR_COSTS=$c timeout 90 python tools/pipeline_parts.py 2>&1 | grep lanes | tee -a gpurun_out/${T}_costs.txt
done
H2_BM_TRACE=1 python -m h2gcn_b200.build --force > /dev/null 2>&1
H2_PAIR_COSTS=${TRACE_COSTS:-8,12,6} timeout 120 python tools/dbg_pair.py i8x3 2>&1 | head -34 | tee gpurun_out/${T}_trace_pair_i8x3.txt
python -m h2gcn_b200.build --force > /dev/null 2>&1
