#!/bin/bash
# Quick GPU check of the tree as it stands: smoke(), the GPU tests, a small bench (plumbing + parity block).
T=${1:-chk}
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/${T}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/${T}_pytest.log
python bench.py --ref-bases 4000000 --reads 300000 --steps 2 --warmup 1 --cpu-seconds 2 > gpurun_out/${T}_small.json 2> gpurun_out/${T}_small.err; echo "small bench exit=$?"
tail -2 gpurun_out/${T}_small.err; head -c 400 gpurun_out/${T}_small.json; echo
