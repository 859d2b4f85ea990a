set -x
python -m pytest tests/test_gpu_upsample.py -q -m gpu 2>&1 | tail -5
ncu --set full --import-source on --clock-control none -k regex:"backward_kernel_up|stats_kernel_up|emit_kernel_up" -c 6 -o gpurun_out/r02g_up python tools/upsample_compare.py 3 > gpurun_out/ncu_up.log 2>&1
tail -3 gpurun_out/ncu_up.log
