#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  B200ZK_NTT_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu --log-n 16 > gpurun_out/nttv_$v.json 2> gpurun_out/nttv_$v.err || echo FAILED $v
done
