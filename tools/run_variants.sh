#!/bin/bash
# usage: tools/run_variants.sh "<label>:<bench args>" ...   -> gpurun_out/sweep_<label>.json
mkdir -p gpurun_out
for spec in "$@"; do
  label="${spec%%:*}"; args="${spec#*:}"
  python bench.py --steps 2 --warmup 3 --no-cpu --no-ntt $args > gpurun_out/sweep_$label.json 2> gpurun_out/sweep_$label.err || echo "FAILED $label"
done
