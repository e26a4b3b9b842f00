#!/bin/bash
T=${1:-r2l}
mkdir -p gpurun_out
python tools/gather_ceiling.py 667 > gpurun_out/${T}_gather_ceiling.jsonl 2> gpurun_out/${T}_gather_ceiling.err; echo "ceiling exit=$?"; cat gpurun_out/${T}_gather_ceiling.jsonl
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab paired GSM_SWEEP_UNPAIRED=0
ab unpaired GSM_SWEEP_UNPAIRED=1
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
