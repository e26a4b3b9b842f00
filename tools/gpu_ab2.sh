#!/bin/bash
# A/B of kernel variants + GPU tests + the full default bench
T=${1:-r2g}
mkdir -p gpurun_out
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab sb8 GSM_SELECT_BLOCKS=8
ab sb6 GSM_SELECT_BLOCKS=6
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench exit=$?"
tail -4 gpurun_out/${T}_bench_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/'+__import__('sys').argv[1] if False else 'gpurun_out/r2g_bench_c4.json'))
print({k:d[k] for k in ('value','ms_per_step','parity','gpu_launches')})
print(d['roofline_methods']); print({k:(v['reads_per_s'],v['ms_per_step']) for k,v in d['methods'].items()})
print(d['e2e']['value'], d['e2e'].get('ascii_input',{}).get('value'), d['e2e'].get('from_fastq'))
PY
