#!/bin/bash
# A/B of kernel variants (sweep block counts, selection variants, op statistics) + GPU tests
T=${1:-r2e}
mkdir -p gpurun_out
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab stats GSM_SWEEP_STATS=1 GSM_SELECT_VARIANT=0
ab mb7 GSM_SWEEP_BLOCKS=7 GSM_SELECT_VARIANT=1
ab mb8 GSM_SWEEP_BLOCKS=8 GSM_SELECT_VARIANT=2
ab v3 GSM_SELECT_VARIANT=3
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
