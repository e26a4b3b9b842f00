#!/bin/bash
T=${1:-r2k}
mkdir -p gpurun_out
python tools/gather_ceiling.py 667,4300 > gpurun_out/${T}_gather_ceiling.jsonl 2> gpurun_out/${T}_gather_ceiling.err; echo "ceiling exit=$?"; cat gpurun_out/${T}_gather_ceiling.jsonl
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --skip-rmi --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab pf0 GSM_SWEEP_PF=0
ab pf64 GSM_SWEEP_PF=64
ab pf128 GSM_SWEEP_PF=128
