#!/bin/bash
# usage: tools/variants.sh ENVVAR "v1 v2 ..." kernel-regex  -> per-variant kernel durations (ncu launch list, one step)
VAR=$1; VALS=$2; KRE=$3
for v in $VALS; do
  env $VAR=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:$KRE --csv --log-file gpurun_out/var_${VAR}_$v.csv python tools/prof_step.py > /dev/null 2>&1
  echo "$VAR=$v $(grep -E "$KRE" gpurun_out/var_${VAR}_$v.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tr '\n' ' ')"
done
