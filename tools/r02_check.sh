#!/bin/bash
# r02: correctness of the fused pack + CTA-pair kernel, then parts timing.   tools/r02_check.sh [pytest -k expression]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="${1:-tensor_core or int8 or row_wise or wide_rounds or non_finite or beyond_2_17 or workspace or zero_copy or auto_mode or random_csr}"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" 2>&1 | tail -25 | tee gpurun_out/r02_check_pytest.txt
for s in i8x3 i8x2; do
  echo "== parts $s"; timeout 120 python tools/time_parts.py $s 2>&1 | tee gpurun_out/r02_parts_$s.txt
done
