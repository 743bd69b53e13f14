#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in ur_gripper ur_admittance basic_env r2d2_maze ur_high_5; do
  timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/f_$cfg.json 2> gpurun_out/f_$cfg.err; echo "$cfg rc=$?"; tail -c 900 gpurun_out/f_$cfg.json; echo
done
