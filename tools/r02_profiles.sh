#!/bin/bash
# r02 evidence: launch list of the bench command, per-kernel traffic of a round, full captures of the pair / pack / dense kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02_launches_bench.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__cycles_active.avg,sm__inst_executed_pipe_tensor.sum --csv --log-file gpurun_out/r02_round_traffic.csv python tools/one_round.py i8x3 128 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:bm_pair -s 3 -c 2 -o gpurun_out/r02_pair -f python tools/one_round.py i8x3 128 6 > /dev/null 2>&1
$NCU --set full -k regex:bm_pack_i8 -s 3 -c 1 -o gpurun_out/r02_pack -f python tools/one_round.py i8x3 128 6 > /dev/null 2>&1
$NCU --set full -k regex:fused_hops_gather -s 3 -c 1 -o gpurun_out/r02_gather -f python tools/one_round.py i8x3 128 6 > /dev/null 2>&1
cat > /tmp/dense_once.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from h2gcn_b200 import ops
a = torch.randn(2708, 448, device="cuda"); w = torch.randn(448, 7, device="cuda")
for _ in range(6):
    ops.dense(a, w); ops.dense(a, w, mode="simt")
torch.cuda.synchronize()
PY
$NCU --set full -k regex:dense -s 6 -c 2 -o gpurun_out/r02_dense -f python /tmp/dense_once.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
