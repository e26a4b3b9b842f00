#!/bin/bash
T=${1:-r2o}
mkdir -p gpurun_out
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab cur GSM_X=0
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
python tools/profile_step.py --method rmi --reads 4000000 --steps 1 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_select" -c 1 -o gpurun_out/${T}_rmi_sel python tools/profile_step.py --method rmi --reads 4000000 --steps 1 > gpurun_out/${T}_ncu1.log 2>&1; echo "ncu rmi exit=$?"
python tools/profile_step.py --method bwa --reads 4000000 --steps 1 > gpurun_out/${T}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_select" -c 1 -o gpurun_out/${T}_bwa_sel python tools/profile_step.py --method bwa --reads 4000000 --steps 1 > gpurun_out/${T}_ncu2.log 2>&1; echo "ncu bwa exit=$?"
