#!/bin/bash
# parity tests, then the A/B harness with the library's defaults (and named environment variants)
T=${1:-r2u}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; python - gpurun_out/${T}_ab_$n.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:d[k] for k in d if k.startswith('ms_select') or k.startswith('sha') or k=='ms_sweep' or k.startswith('records')})
PY
tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab def GSM_X=0

