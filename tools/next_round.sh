#!/bin/bash
# First measurements of the next round (one gpurun call, ~2 GPU-minutes):  tools/next_round.sh
#  1. tools/umma_i8_probe.cu part 3: completion latency of tcgen05.st with / without MMAs in flight
#  2. the CTA-pair kernel (H2_BM_PAIR=1) with 2 / 4 producer groups per CTA and 1 / 2 units per A hand-over: parity check +
#     tensor-hop time + round time (each variant is a rebuild: ~40 s; the parity check has a 60 s timeout against hangs)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_i8_probe tools/umma_i8_probe.cu 2>/dev/null
timeout 90 /tmp/umma_i8_probe 2>&1 | tail -6 | tee gpurun_out/next_sttm_latency.log
for CFG in "2 1 8" "4 1 8" "2 2 8" "4 2 12"; do
  set -- $CFG; G=$1; U=$2; B=$3
  H2_EXTRA_NVCC_FLAGS="-DH2_BM_PAIR_GROUPS=$G -DH2_BM_PAIR_UNITS=$U -DH2_BM_PAIR_B_STAGES=$B" python -m h2gcn_b200.build --force > /dev/null 2>&1
  echo "== pair kernel: $G producer groups per CTA, $U unit(s) per hand-over, $B B stages"
  H2_BM_PAIR=1 timeout 60 python tools/pair_check.py 2>&1 | tail -2
  H2_BM_PAIR=1 timeout 60 python tools/time_parts.py i8x2 2>&1 | grep -E "tensor only|full round \(flush"
  H2_BM_PAIR=1 timeout 90 python bench.py --steps 300 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; l=json.loads(sys.stdin.read()); print('bench us/round', round(l['ms_per_step']*1e3,2))"
done
python -m h2gcn_b200.build --force > /dev/null 2>&1
