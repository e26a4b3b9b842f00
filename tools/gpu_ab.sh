#!/bin/bash
# One GPU session of A/B runs: (optionally) the GPU tests first, then tools/sweep_ab.py once per variant.  A variant is a
# quoted list of environment assignments; record checksums (sha_*) must be identical across variants.
#   usage: tools/gpu_ab.sh TAG [--tests] "GSM_SELECT_OPT=0" "GSM_SELECT_OPT=3" "GSM_SWEEP_BLOCKS=6 GSM_BWA_PICKS=0" ...
T=$1; shift
mkdir -p gpurun_out
if [ "$1" == "--tests" ]; then
  shift
  python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
fi
i=0
for v in "$@"; do
  i=$((i+1))
  env $v python tools/sweep_ab.py --tag "$v" > gpurun_out/${T}_ab_$i.json 2> gpurun_out/${T}_ab_$i.err; echo "[$v] exit=$?"
  python - gpurun_out/${T}_ab_$i.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in d if k.startswith(("ms_select", "sha", "records")) or k == "ms_sweep"})
PY
  tail -2 gpurun_out/${T}_ab_$i.err | cut -c1-300
done
