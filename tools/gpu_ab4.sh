#!/bin/bash
# parity tests, then the A/B harness under several selection-kernel variants
T=${1:-r2p}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab b8 GSM_SELECT_BLOCKS=8
ab b7 GSM_SELECT_BLOCKS=7
ab b6 GSM_SELECT_BLOCKS=6
