#!/bin/bash
# Round-end evidence on one B200: bench lines for every config (with the CPU baseline), the reference arm, the ncu launch
# list of the default bench command, one ncu --set full capture of the step kernel per config, the rollout parity report.
mkdir -p gpurun_out/g
O=gpurun_out/g
python bench.py > $O/bench_r2d2_maze.json 2> $O/bench_r2d2_maze.err; echo "default bench rc=$?"; tail -c 300 $O/bench_r2d2_maze.json; echo
python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "reference arm rc=$?"; cat $O/bench_reference_arm.json | cut -c1-300
for cfg in ur_high_5 ur_high_5_randomised from_the_readme drone_pilot basic_env ur_admittance ur_gripper; do
  timeout 600 python bench.py --config $cfg > $O/bench_$cfg.json 2> $O/bench_$cfg.err; echo "$cfg rc=$?"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2d2_maze.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
for spec in r2d2_maze:4096:60 ur_high_5:8192:10 basic_env:4096:60 ur_gripper:4096:30; do
  IFS=: read cfg n skip <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dg_step_kernel -s $skip -c 1 -f -o $O/step_$cfg python tools/profile_cmd.py $cfg 0 $n $((skip+2)) > $O/ncu_full_$cfg.log 2>&1; echo "ncu full $cfg rc=$?"
done
timeout 1200 python tools/parity_report.py --configs ur_high_5,ur_high_5_randomised,r2d2_maze,from_the_readme,basic_env,ur_admittance,ur_gripper > $O/parity_report.json 2> $O/parity_report.err; echo "parity report rc=$?"
ls -la $O | head -40
