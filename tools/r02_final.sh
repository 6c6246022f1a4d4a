#!/bin/bash
# r02 final evidence, one gpurun call:  tools/r02_final.sh <tag>   (writes gpurun_out/<tag>_*)
#   full GPU test suite, smoke, bench (ours + reference arm), launch list, per-kernel traffic of a round, full ncu captures
#   of the pair / gather / pack kernels, per-CTA trace of the pair kernel
cd "$(dirname "$0")/.."
T=${1:-r02z}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/${T}_gputests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/${T}_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > $O/${T}_clocks.csv &
SMI=$!
timeout 300 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 200 python bench.py --steps 20 --warmup 3 > $O/${T}_bench_steps20.json 2>> $O/${T}_bench.err
kill $SMI
timeout 120 python bench.py --impl reference --steps 20 --warmup 3 > $O/${T}_bench_reference_arm.json 2>> $O/${T}_bench.err
timeout 300 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file $O/${T}_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary > $O/${T}_launches_bench.log 2>&1
timeout 200 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__cycles_active.avg,sm__inst_executed_pipe_tensor.sum,smsp__inst_executed.sum -k regex:"bm_|fused_hops" --csv --log-file $O/${T}_round_traffic.csv python tools/one_round.py i8x3 128 3 > /dev/null 2>&1
timeout 200 $NCU --set full --import-source on -k regex:bm_pair -s 3 -c 1 -o $O/${T}_pair -f python tools/one_round.py i8x3 128 6 > /dev/null 2>&1
timeout 200 $NCU --set full --import-source on -k regex:fused_hops_gather -s 3 -c 1 -o $O/${T}_gather -f python tools/one_round.py i8x3 128 6 > /dev/null 2>&1
timeout 200 $NCU --set full -k regex:bm_pack_i8 -s 3 -c 1 -o $O/${T}_pack -f python tools/one_round.py i8x3 128 6 > /dev/null 2>&1
for k in pair gather pack; do
  ncu -i $O/${T}_$k.ncu-rep --page raw --csv > $O/${T}_${k}_ncu_raw.csv 2>/dev/null
done
ncu -i $O/${T}_pair.ncu-rep --page source --csv > $O/${T}_pair_ncu_source.csv 2>/dev/null
ncu -i $O/${T}_gather.ncu-rep --page source --csv > $O/${T}_gather_ncu_source.csv 2>/dev/null
H2_BM_TRACE=1 python -m h2gcn_b200.build --force > /dev/null 2>&1
H2_TRACE_DUMP=$O/${T}_pair_ctas.csv timeout 120 python tools/dbg_pair.py i8x3 2>&1 | head -40 > $O/${T}_trace_pair_i8x3.txt
python -m h2gcn_b200.build --force > /dev/null 2>&1
cut -c1-300 $O/${T}_bench.json; echo; cut -c1-200 $O/${T}_bench_steps20.json; echo; cut -c1-300 $O/${T}_bench_reference_arm.json; echo; ls -la $O | grep ${T}_ | awk '{print $5, $9}'
