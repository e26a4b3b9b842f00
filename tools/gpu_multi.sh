#!/bin/bash
# Multi-GPU session: bench.py under torchrun at N ranks (small plumbing run first, then the default config).
N=${1:-2}; T=${2:-r2n}; shift 2
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run --ref-bases 4000000 --reads 400000 --steps 2 --warmup 1 > gpurun_out/${T}_small_n$N.json 2> gpurun_out/${T}_small_n$N.err; echo "small exit=$?"
grep -v "^\[W\|NCCL INFO" gpurun_out/${T}_small_n$N.err | tail -12
head -c 1500 gpurun_out/${T}_small_n$N.json; echo
run "$@" > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench exit=$?"
grep -v "^\[W\|NCCL INFO" gpurun_out/${T}_bench_n$N.err | tail -12
head -c 600 gpurun_out/${T}_bench_n$N.json; echo
