#!/bin/bash
# Final single-GPU session: smoke, the default bench (C4), the C3 bench, then (after those exited) the ncu launch list of a
# short bench run and one full capture of the dominant kernel.  Logs under gpurun_out/.
T=${1:-r2x}
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; echo "bench c4 exit=$?"
tail -4 gpurun_out/${T}_bench_c4.err; head -c 1500 gpurun_out/${T}_bench_c4.json; echo
python bench.py --config c3 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; echo "bench c3 exit=$?"
tail -3 gpurun_out/${T}_bench_c3.err; head -c 800 gpurun_out/${T}_bench_c3.json; echo
python bench.py --reads 4000000 --steps 2 --warmup 1 --skip-extras --cpu-seconds 0 > gpurun_out/${T}_short.json 2> gpurun_out/${T}_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --reads 4000000 --steps 2 --warmup 1 --skip-extras --cpu-seconds 0 > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list exit=$?"
python tools/profile_step.py --method bwa --reads 4000000 --steps 1 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_sweep1" -c 1 -o gpurun_out/${T}_sweep python tools/profile_step.py --method bwa --reads 4000000 --steps 1 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full exit=$?"
