#!/bin/bash
# A/B: L2 persistence window on the LUT / the RMI parameters; RMI at 6 CTAs per SM
T=${1:-r2z}
mkdir -p gpurun_out
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; python - gpurun_out/${T}_ab_$n.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:d[k] for k in d if k.startswith('ms_select') or k=='ms_sweep' or k.startswith('sha')})
PY
tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab def GSM_X=0
ab persist AB_PERSIST=1
ab b6 GSM_SELECT_BLOCKS=6
