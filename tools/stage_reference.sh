#!/bin/bash
# Stage the python packages of the reference that the integration run imports (losses, models, utils, managers, configs)
# under baseline/_ref/ -- git-ignored, but shipped to the GPU box by gpurun, which has no /root/reference.
# Nothing under baseline/_ref is part of this repository's sources; tools/train_step_ocrnet.py and
# tests/test_gpu_reference_integration.py import it as the UNMODIFIED reference.
set -e
SRC=${1:-/root/reference}
DST=$(dirname "$0")/../baseline/_ref
mkdir -p "$DST"
for d in losses models utils managers configs datasets; do
  rm -rf "$DST/$d"
  cp -r "$SRC/$d" "$DST/$d"
done
cp "$SRC/main.py" "$DST/" 2>/dev/null || true
find "$DST" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
du -sh "$DST"
