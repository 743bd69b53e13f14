python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu5.log 2>&1; tail -3 gpurun_out/pytest_gpu5.log
python tools/gpu_probe3.py ur_high_5:2:0 ur_high_5:4:0 ur_high_5:8:0 ur_high_5:4:0:2 r2d2_maze:2:0 r2d2_maze:4:0 r2d2_maze:8:0 r2d2_maze:8:0:3 from_the_readme:4:0 from_the_readme:8:0 from_the_readme:4:0:3 drone_pilot:4:0 basic_env:4:0 ur_high_5:4:0:0:65536 > gpurun_out/probe9.log 2>&1
cat gpurun_out/probe9.log
python tools/profile_cmd.py ur_high_5 4 8192 3 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dg_step -s 3 -c 1 -o gpurun_out/prof_r1g_ur python tools/profile_cmd.py ur_high_5 4 8192 3 > gpurun_out/ncu4.log 2>&1
tail -1 gpurun_out/ncu4.log
