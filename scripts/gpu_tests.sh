#!/bin/bash
# Runs every GPU test file under its own timeout, logging to gpurun_out/ (partial results survive).
mkdir -p gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))"
for f in "$@"; do
  n=$(basename $f .py)
  timeout 300 python -m pytest $f -m gpu -x -q -s --durations=5 > gpurun_out/$n.log 2>&1
  echo "== $n rc=$?"; tail -15 gpurun_out/$n.log
done
