#!/bin/bash
# usage: ncu_generic.sh <workload> <kernel regex> <out name> [skip] [env...]
out=gpurun_out; mkdir -p $out
w=$1; k=$2; name=$3; skip=${4:-12}
ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f \
  -o $out/$name python tools/quick_step.py $w --steps 4 > $out/$name.log 2>&1
tail -2 $out/$name.log
