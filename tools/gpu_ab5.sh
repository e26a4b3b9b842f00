#!/bin/bash
# parity tests, then the A/B harness under selection-kernel variants (GSM_SELECT_BLOCKS x GSM_SELECT_OPT); records must keep their SHA
T=${1:-r2s}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; python - gpurun_out/${T}_ab_$n.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:d[k] for k in d if k.startswith('ms_select') or k.startswith('sha') or k=='ms_sweep'})
PY
tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab b7o0 GSM_SELECT_BLOCKS=7 GSM_SELECT_OPT=0
ab b7o1 GSM_SELECT_BLOCKS=7 GSM_SELECT_OPT=1
ab b7o2 GSM_SELECT_BLOCKS=7 GSM_SELECT_OPT=2
ab b7o3 GSM_SELECT_BLOCKS=7 GSM_SELECT_OPT=3
ab b6o3 GSM_SELECT_BLOCKS=6 GSM_SELECT_OPT=3
ab b8o3 GSM_SELECT_BLOCKS=8 GSM_SELECT_OPT=3
