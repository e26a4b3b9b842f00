#!/bin/bash
# parity tests, then the A/B harness with the library's defaults
T=${1:-r2u}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
python tools/sweep_ab.py --tag defaults > gpurun_out/${T}_ab_def.json 2> gpurun_out/${T}_ab_def.err; echo "exit=$?"; cat gpurun_out/${T}_ab_def.json; tail -2 gpurun_out/${T}_ab_def.err | cut -c1-300
