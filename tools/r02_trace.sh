#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
H2_BM_TRACE=1 python -m h2gcn_b200.build --force > /dev/null 2>&1
for s in ${@:-i8x3 i8x2}; do
  timeout 120 python tools/dbg_pair.py $s 2>&1 | tee gpurun_out/r02_trace_pair_$s.txt
done
python -m h2gcn_b200.build --force > /dev/null 2>&1
