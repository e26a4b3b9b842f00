#!/bin/bash
# 8-GPU session: configs[3] (strong scaling, 50 M reads sharded over 8 GPUs)
T=${1:-r2n}
mkdir -p gpurun_out
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@"; }
run 8 --steps 5 --warmup 3 --skip-extras > gpurun_out/${T}_bench_c4_n8.json 2> gpurun_out/${T}_bench_c4_n8.err; echo "c4 n8 exit=$?"
grep -v "^\[W\|NCCL INFO" gpurun_out/${T}_bench_c4_n8.err | tail -4; head -c 300 gpurun_out/${T}_bench_c4_n8.json; echo
