#!/bin/bash
# traces of the pair kernel under several schedule settings (library built with H2_BM_TRACE), then per-part times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-sweep}
H2_BM_TRACE=1 python -m h2gcn_b200.build --force > /dev/null 2>&1
i=0
for c in ${COSTS:-"8,12,6"}; do
  i=$((i+1))
  echo "== trace H2_PAIR_COSTS=$c" | tee -a gpurun_out/${T}_traces.txt
  H2_TRACE_DUMP=gpurun_out/${T}_ctas_$i.csv H2_PAIR_COSTS=$c timeout 120 python tools/dbg_pair.py i8x3 2>&1 | head -30 | tee -a gpurun_out/${T}_traces.txt
done
echo "== trace no finisher" | tee -a gpurun_out/${T}_traces.txt
H2_PAIR_NO_FINISHER=1 H2_TRACE_DUMP=gpurun_out/${T}_ctas_nofin.csv H2_PAIR_COSTS=8,12,6 timeout 120 python tools/dbg_pair.py i8x3 2>&1 | head -30 | tee -a gpurun_out/${T}_traces.txt
python -m h2gcn_b200.build --force > /dev/null 2>&1
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor or int8 or i8 or pair or bitmap or round" 2>&1 | tail -2 | tee gpurun_out/${T}_tests.txt
for c in ${COSTS:-"8,12,6"}; do
  echo "== H2_PAIR_COSTS=$c" | tee -a gpurun_out/${T}_costs.txt
  H2_PAIR_COSTS=$c timeout 90 python tools/pipeline_parts.py 2>&1 | grep "lanes 2" | tee -a gpurun_out/${T}_costs.txt
done
