#!/bin/bash
# One GPU-box visit: the -m gpu suite (per-test time limit, first failure stops it), then a short bench.  Everything is logged
# under gpurun_out/ so that a command cut off by the box's limit still leaves its output behind.
# usage: tools/gpu_ci.sh <tag> [pytest args...]
tag=${1:-ci}; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail 8 --timeout 420 -p no:cacheprovider "$@" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
