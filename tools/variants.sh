for v in Z1 Z2 Z3; do
  B200SEG_STATS_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:stats_kernel_ --csv --log-file gpurun_out/var_$v.csv python tools/prof_step.py > /dev/null 2>&1
done
