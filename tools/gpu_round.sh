#!/bin/bash
# One GPU box visit: the GPU test-suite, then bench lines.  Everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_round.sh <tag> [pytest-args...]
set -u
TAG=${1:-run}
shift || true
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt; free -g | head -2 >> $O/${TAG}_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --durations=15 "$@" > $O/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
