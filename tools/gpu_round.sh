#!/bin/bash
# One GPU session: smoke, a small bench (plumbing check), the GPU tests, the full default bench.  Logs under gpurun_out/.
T=${1:-r2a}
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit=$?"
python bench.py --ref-bases 4000000 --reads 300000 --steps 2 --warmup 1 --cpu-seconds 2 > gpurun_out/${T}_small.json 2> gpurun_out/${T}_small.err; echo "small bench exit=$?"
tail -3 gpurun_out/${T}_small.err
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"
tail -5 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench exit=$?"
tail -5 gpurun_out/${T}_bench_c4.err
head -c 3000 gpurun_out/${T}_bench_c4.json
