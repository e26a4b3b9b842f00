#!/bin/bash
# A/B + GPU tests + the default bench (C4) + C3 + launch list + one full ncu capture of the sweep kernel
T=${1:-r2j}
mkdir -p gpurun_out
ab() { n=$1; shift; env "$@" python tools/sweep_ab.py --tag "$*" > gpurun_out/${T}_ab_$n.json 2> gpurun_out/${T}_ab_$n.err; echo "[$*] exit=$?"; cat gpurun_out/${T}_ab_$n.json; tail -2 gpurun_out/${T}_ab_$n.err | cut -c1-300; }
ab final GSM_SELECT_BLOCKS=8
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench c4 exit=$?"; tail -2 gpurun_out/${T}_bench_c4.err
python bench.py --config c3 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; echo "bench c3 exit=$?"; tail -2 gpurun_out/${T}_bench_c3.err
python tools/gather_ceiling.py 64,667,4300 > gpurun_out/${T}_gather_ceiling.jsonl 2> gpurun_out/${T}_gather_ceiling.err; echo "ceiling exit=$?"; tail -3 gpurun_out/${T}_gather_ceiling.jsonl
python tools/profile_step.py --method bwa --reads 4000000 --steps 1 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/profile_step.py --method bwa --reads 4000000 --steps 1 > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches exit=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_sweep1" -s 1 -c 1 -o gpurun_out/${T}_sweep python tools/profile_step.py --method bwa --reads 4000000 --steps 1 > gpurun_out/${T}_ncu.log 2>&1; echo "ncu full exit=$?"
