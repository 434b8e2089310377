#!/bin/bash
# The round's ncu evidence, on one GPU: launch list of the bench command, then one full capture each of the two kernels the
# roofline numbers are about.  Every ncu pass runs the very command that has just exited 0 without ncu.
# usage: tools/profile_round.sh <tag>     (outputs under gpurun_out/)
tag=${1:-r02}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-circuits"
mkdir -p gpurun_out
$CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel_v6 -s 4 -c 1 -f -o gpurun_out/${tag}_msm_acc $CMD > gpurun_out/${tag}_ncu_msm.log 2>&1
echo "msm capture rc=$?"
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel_lb0 -s 6 -c 2 -f -o gpurun_out/${tag}_ntt $CMD > gpurun_out/${tag}_ncu_ntt.log 2>&1
echo "ntt capture rc=$?"
for r in msm_acc ntt; do
  if [ -f gpurun_out/${tag}_${r}.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${tag}_${r}.ncu-rep gpurun_out/${tag}_ncu_${r}.json "${tag} $r: $CMD"
    ncu -i gpurun_out/${tag}_${r}.ncu-rep --page source --csv > gpurun_out/${tag}_${r}_source.csv 2>/dev/null
    ls -la gpurun_out/${tag}_${r}.ncu-rep
    # the report itself only travels back if it is small
    [ $(stat -c %s gpurun_out/${tag}_${r}.ncu-rep) -gt 25000000 ] && rm -f gpurun_out/${tag}_${r}.ncu-rep
  fi
done
python tools/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches.txt 2>&1
tail -5 gpurun_out/${tag}_launches.txt
