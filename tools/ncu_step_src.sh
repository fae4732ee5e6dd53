#!/bin/bash
# usage: tools/ncu_step_src.sh <variant> <ept> <envs>   (scratch: source-counter ncu capture of maze_step)
v=$1; ept=$2; envs=$3
export MAZE_B200_LIB=/root/repo/ab_libs/lib_$v.so MAZE_STEP_EPT=$ept
python tools/perf_step.py $envs > gpurun_out/plain_${v}_${ept}.log 2>&1 && \
ncu --section SourceCounters --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --import-source on --cache-control none --clock-control none -k regex:maze_step -s 400 -c 1 -o gpurun_out/src_${v}_${ept} -f python tools/perf_step.py $envs > gpurun_out/ncusrc_${v}_${ept}.log 2>&1
tail -n 2 gpurun_out/ncusrc_${v}_${ept}.log
