#!/bin/bash
# End-of-round evidence, one gpurun call:  tools/final_evidence.sh <tag>   (writes gpurun_out/<tag>_*)
cd "$(dirname "$0")/.."
T=${1:-r01f}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/${T}_clocks.csv &
SMI=$!
timeout 240 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
kill $SMI
timeout 120 python bench.py --impl reference --steps 20 --warmup 3 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/${T}_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__cycles_active.avg --clock-control none -k regex:"bm_|fused_hops" -s 45 -c 10 --csv --log-file $O/${T}_round_traffic.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:bm_mma -s 4 -c 1 -o $O/${T}_bm_mma python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
tail -c 400 $O/${T}_bench.json; echo; tail -2 $O/${T}_smoke.log; ls -la $O | tail -12
