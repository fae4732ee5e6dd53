#!/bin/bash
# usage: tools/ncu_step.sh <layout> <envs> [lib-variant]   (scratch: few-metric ncu capture of steady-state maze_step)
lay=$1; envs=$2; v=${3:-}
[ -n "$v" ] && export MAZE_B200_LIB=/root/repo/ab_libs/lib_$v.so
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"
ncu --metrics $M --cache-control none --clock-control none -k regex:maze_step -s 900 -c 2 --csv --log-file gpurun_out/ncu_${lay}_${envs}_${v}.csv python tools/perf_step.py $lay $envs > gpurun_out/ncu_${lay}_${envs}_${v}.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_${lay}_${envs}_${v}.csv")) if len(r)>10]
h=rows[0]; i_n=h.index("Metric Name"); i_v=h.index("Metric Value"); i_id=h.index("ID")
d={}
for r in rows[1:]:
    if r[i_id]==rows[1][i_id]: d[r[i_n]]=r[i_v]
print("$lay $envs $v", {k.split('.')[0].replace('__','_')[-28:]:v for k,v in d.items()})
PY
