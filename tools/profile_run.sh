#!/bin/bash
# ncu evidence of one training step (1 GPU): launch list + `--set full` of every launch of one step (exported to CSV on the box:
# the .ncu-rep of 52 launches is > 100 MB) + source-level capture of the dominant kernel.  usage: tools/profile_run.sh <tag>
tag=${1:-r2p}
B="python bench.py --launch eager --no-cpu-baseline --no-torch-gpu --no-train-py"
ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv $B --steps 2 --warmup 3 > /dev/null 2> gpurun_out/${tag}_ncu_a.log
ncu --set full --clock-control none -s 230 -c 52 -o /tmp/${tag}_full $B --steps 2 --warmup 3 > /dev/null 2> gpurun_out/${tag}_ncu_b.log
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>> gpurun_out/${tag}_ncu_b.log
ncu --set full --clock-control none --import-source on -k regex:score_bwd_mma -s 4 -c 1 -o gpurun_out/${tag}_score_bwd $B --steps 2 --warmup 3 > /dev/null 2> gpurun_out/${tag}_ncu_c.log
ls -la gpurun_out/${tag}_*
