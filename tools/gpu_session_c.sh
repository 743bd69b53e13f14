#!/bin/bash
# ncu --set full captures of the step kernel late in a short r2d2_maze run (launch 60: R2D2s on the ground / at walls)
mkdir -p gpurun_out
for S in "$@"; do
  DG_SOLVER=$S DG_RS_ASHARED=0 timeout 300 python tools/profile_cmd.py r2d2_maze 0 4096 62 || exit 1
  DG_SOLVER=$S DG_RS_ASHARED=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dg_step_kernel -s 60 -c 1 -f -o gpurun_out/step_r2d2_solver$S python tools/profile_cmd.py r2d2_maze 0 4096 62 > gpurun_out/ncu_c$S.log 2>&1
  tail -2 gpurun_out/ncu_c$S.log
done
ls -la gpurun_out/*.ncu-rep
