for v in 0 1 2 3 4 5 6; do
  B200SEG_BWD_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:backward --csv --log-file gpurun_out/var_$v.csv python tools/prof_step.py > /dev/null 2>&1
  echo "variant=$v $(grep backward gpurun_out/var_$v.csv | awk -F'","' '{print $5, $NF}' | cut -c1-100)"
done
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
