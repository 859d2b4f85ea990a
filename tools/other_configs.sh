#!/bin/bash
# Kernel times (ncu launch list, one step) of the non-headline configurations: per-image C=17 (BASELINE configs[1]),
# trained-like blocky logits (D2 of SURVEY 8d), C=8.
for cfg in "--classes 17 --per-image" "--classes 25 --blocky" "--classes 8" "--classes 25 --per-image"; do
  tag=$(echo $cfg | tr -d ' -')
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/cfg_$tag.csv python tools/prof_step.py $cfg > /dev/null 2>&1
  echo "== $cfg"; python tools/launch_times.py gpurun_out/cfg_$tag.csv | grep -v "at::" | awk '{printf "%s %s %s | ", $1, $3, $4} END {print ""}'
done
