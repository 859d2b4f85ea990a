#!/bin/bash
# In-stream per-stage times (CUDA events between the kernels of a running step), no CPU baseline / e2e legs.
python bench.py --no-cpu-baseline --no-e2e --steps 40 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('ms', round(d['ms_per_step'],4), {k: round(v*1000,1) for k,v in d['stage_ms'].items()})"
