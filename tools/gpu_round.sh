#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu full capture of our kernels.
# usage: tools/gpu_round.sh <tag> [skip_tests]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/${TAG}_launches.csv python tools/prof_step.py > $OUT/${TAG}_launch.log 2>&1
python tools/launch_times.py $OUT/${TAG}_launches.csv | grep -v "at::" 
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'stats_kernel|finalize_decide|emit_kernel|sort_prepare|hyb_|sort_fallback|backward_kernel|metrics_kernel' \
  -o $OUT/${TAG}_full -f python tools/prof_step.py > $OUT/${TAG}_full.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_full.ncu-rep
