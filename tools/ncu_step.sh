#!/bin/bash
# usage: tools/ncu_step.sh <variant> <ept> <envs>   (scratch: few-metric ncu capture of maze_step)
v=$1; ept=$2; envs=$3
export MAZE_B200_LIB=/root/repo/ab_libs/lib_$v.so MAZE_STEP_EPT=$ept
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__d_sectors_fill_device.sum,lts__t_sectors_srcnode_gpc_evict_first.sum"
python tools/perf_step.py $envs > gpurun_out/plain_${v}_${ept}.log 2>&1 && \
ncu --metrics $M --cache-control none --clock-control none -k regex:maze_step -s 400 -c 3 --csv --log-file gpurun_out/ncu_${v}_${ept}.csv python tools/perf_step.py $envs > gpurun_out/ncu_${v}_${ept}.log 2>&1
tail -n 2 gpurun_out/ncu_${v}_${ept}.log
