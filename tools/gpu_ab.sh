#!/bin/bash
# A/B of the sweep variants + GPU tests under the variant + ncu captures of the LUT step's kernels
T=${1:-r2c}
mkdir -p gpurun_out
ab() { env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$N.json 2> gpurun_out/${T}_ab_$N.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$N.json; tail -2 gpurun_out/${T}_ab_$N.err | cut -c1-300; }
N=pair ab GSM_SWEEP_LPR=2
N=lane ab GSM_SWEEP_LPR=1 GSM_SWEEP_UNIQ=0
N=lane_uniq ab GSM_SWEEP_LPR=1 GSM_SWEEP_UNIQ=1
GSM_SWEEP_LPR=1 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_lpr1.log 2>&1; echo "pytest lpr1 uniq exit=$?"; tail -4 gpurun_out/${T}_pytest_lpr1.log
export GSM_SWEEP_LPR=1
python tools/profile_step.py --method lut --reads 1000000 --steps 1 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_select_seeded|k_sweep" -c 2 -o gpurun_out/${T}_lut python tools/profile_step.py --method lut --reads 1000000 --steps 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu exit=$?"
tail -3 gpurun_out/${T}_ncu.log
