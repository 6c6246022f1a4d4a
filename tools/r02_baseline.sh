#!/bin/bash
# r02a: where the round-1 kernels stand with the fp32-equivalent arithmetic (i8x3) — parts, traces, STTM latency probe
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r02a_smi.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_i8_probe tools/umma_i8_probe.cu 2>/dev/null
timeout 90 /tmp/umma_i8_probe 2>&1 | tail -12 | tee gpurun_out/r02a_sttm_latency.log
for s in i8x3 i8x2; do
  echo "== parts $s"; timeout 120 python tools/time_parts.py $s 2>&1 | tee gpurun_out/r02a_parts_$s.txt
done
H2_BM_TRACE=1 python -m h2gcn_b200.build --force > /dev/null 2>&1
for s in i8x3 i8x2; do
  timeout 120 python tools/dbg_run.py $s alone > gpurun_out/r02a_trace_$s.txt 2>&1; head -4 gpurun_out/r02a_trace_$s.txt
done
python -m h2gcn_b200.build --force > /dev/null 2>&1
