#!/bin/bash
mkdir -p gpurun_out/j
for spec in basic_env:4096:60 ur_gripper:4096:30; do
  IFS=: read cfg n skip <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dg_step_kernel -s $skip -c 1 -f -o gpurun_out/j/step_$cfg python tools/profile_cmd.py $cfg 0 $n $((skip+2)) > gpurun_out/j/ncu_full_$cfg.log 2>&1; echo "ncu full $cfg rc=$?"
done
