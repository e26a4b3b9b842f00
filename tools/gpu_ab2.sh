#!/bin/bash
# quick A/B of the lane sweep kernel (+ tests) after a change; ncu capture of the sweep + the RMI select
T=${1:-r2d}
mkdir -p gpurun_out
export GSM_SWEEP_LPR=1
for mb in 0 7 8; do
  GSM_SWEEP_BLOCKS=$mb python tools/sweep_ab.py --tag "$T lane uniq blocks=$mb" > gpurun_out/${T}_ab_mb$mb.json 2> gpurun_out/${T}_ab_mb$mb.err; echo "ab mb=$mb exit=$?"; cat gpurun_out/${T}_ab_mb$mb.json; tail -2 gpurun_out/${T}_ab_mb$mb.err | cut -c1-300
done
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
python tools/profile_step.py --method rmi --reads 1000000 --steps 1 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_select|k_sweep" -c 2 -o gpurun_out/${T}_rmi python tools/profile_step.py --method rmi --reads 1000000 --steps 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit=$?"
tail -2 gpurun_out/${T}_ncu.log
