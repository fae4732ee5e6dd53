#!/bin/bash
# usage (on the GPU box, from the repo root): tools/round_profile.sh <tag>
# The round's evidence in one call: GPU tests, both bench arms, the ncu launch list of the headline command
# and one full steady-state capture of the step kernel.  Everything lands in gpurun_out/ (scratch); copy what
# should be judged into profiles/.
tag=${1:-r01x}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $o/pytest_gpu_$tag.log
python bench.py --impl reference > $o/bench_${tag}_reference.json 2> $o/bench_${tag}_reference.err; echo "reference arm rc=$?"
python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"
cat $o/bench_$tag.json
# launch list of the headline path (no extras): the timed region launches maze_step_kernel only
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_$tag.csv \
    python bench.py --steps 100 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 3 > $o/ncu_bench_$tag.log 2>&1; echo "ncu list rc=$?"
# steady-state capture of the step kernel (launch 1200: episodes desynchronised)
ncu --set full --cache-control none --clock-control none --import-source on -k regex:maze_step_kernel -s 1200 -c 1 -f -o $o/step_$tag \
    python bench.py --steps 1500 --warmup 3 --no-extras --no-cpu-baseline --e2e-steps 3 > $o/ncu_step_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i $o/step_$tag.ncu-rep --page details > $o/step_${tag}_details.txt
ncu -i $o/step_$tag.ncu-rep --page raw --csv > $o/step_${tag}_raw.csv
