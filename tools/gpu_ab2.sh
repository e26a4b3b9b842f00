#!/bin/bash
# A/B of kernel variants (sweep block counts, op statistics) + GPU tests + ncu of the LUT step at 4 M reads
T=${1:-r2f}
mkdir -p gpurun_out
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab mb7 GSM_SWEEP_BLOCKS=0
ab mb8 GSM_SWEEP_BLOCKS=8
ab mb6 GSM_SWEEP_BLOCKS=6
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
python tools/profile_step.py --method lut --reads 4000000 --steps 1 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_select|k_sweep" -c 2 -o gpurun_out/${T}_lut python tools/profile_step.py --method lut --reads 4000000 --steps 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit=$?"
tail -2 gpurun_out/${T}_ncu.log
