#!/bin/bash
# multi-GPU evidence: 2-GPU tests (fp32 + bf16 shards, NCCL vs peer-memory exchange), cfg5 at reduced scale, cfg4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
T=${2:-r02j}
[ -z "$SKIP_TESTS" ] && timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${T}_multi_tests.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
H2_BENCH_CFG5_N=${CFG5_N:-1048576} H2_BENCH_CFG5_E=${CFG5_E:-16777216} timeout 600 $TR bench.py --gpus $N --workload cfg5 --steps ${CFG5_STEPS:-5} --warmup 3 > gpurun_out/${T}_cfg5_${N}gpu.json 2> gpurun_out/${T}_cfg5_${N}gpu.err
tail -2 gpurun_out/${T}_cfg5_${N}gpu.err; cut -c1-300 gpurun_out/${T}_cfg5_${N}gpu.json
if [ -z "$SKIP_CFG4" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_cfg4_${N}gpu.json 2> gpurun_out/${T}_cfg4_${N}gpu.err
tail -2 gpurun_out/${T}_cfg4_${N}gpu.err; cut -c1-300 gpurun_out/${T}_cfg4_${N}gpu.json
fi
