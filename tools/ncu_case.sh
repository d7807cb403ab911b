#!/bin/bash
# usage: tools/ncu_case.sh <case> <kernel-regex>   (on the GPU box; keeps only small CSV extracts of the capture)
c=$1; k=$2; skip=${3:-1}
python tools/prof_r2.py $c > gpurun_out/prof_$c.log 2>&1 || { echo "$c plain run failed"; exit 1; }
ncu --set full --clock-control none -s $skip -c 1 -k regex:"$k" -o /tmp/r2_prof_$c -f python tools/prof_r2.py $c > gpurun_out/ncu_$c.log 2>&1
ncu -i /tmp/r2_prof_$c.ncu-rep --page details --csv > gpurun_out/r2_ncu_${c}_details.csv 2>/dev/null
ncu -i /tmp/r2_prof_$c.ncu-rep --page raw --csv > /tmp/raw_$c.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("/tmp/raw_$c.csv")))
hdr=rows[0]; val=rows[2] if len(rows)>2 else rows[-1]
keep=[i for i,h in enumerate(hdr) if any(s in h for s in ("Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct","launch__registers","launch__grid_size","sm__warps_active.avg.pct","smsp__issue_active.avg.pct","smsp__inst_executed.sum","l1tex__t_sector_hit_rate","lts__t_sector_hit_rate","smsp__average_warp","sm__throughput.avg.pct","launch__occupancy_limit","sm__maximum_warps","achieved_occupancy","l1tex__data_pipe","lsu","warp_issue_stalled","smsp__pcsamp","dram__cycles_active","lts__t_bytes","l1tex__t_bytes","sm__inst_executed_pipe_lsu","tensor","tmem","dmma","pipe_fp64"))]
with open("gpurun_out/r2_ncu_${c}_raw.txt","w") as f:
    for i in keep: f.write(f"{hdr[i]} = {val[i]}\n")
PY
rm -f /tmp/r2_prof_$c.ncu-rep /tmp/raw_$c.csv
echo "$c done"
