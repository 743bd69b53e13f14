python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu6.log 2>&1; tail -3 gpurun_out/pytest_gpu6.log
python tools/gpu_probe3.py ur_high_5:2:0:0 ur_high_5:4:0:0 ur_high_5:8:0:0 r2d2_maze:4:0:0 r2d2_maze:8:0:0 r2d2_maze:16:0:0 r2d2_maze:8:0:2 from_the_readme:4:0:0 from_the_readme:8:0:0 drone_pilot:4:0:0 basic_env:4:0:0 ur_high_5:4:0:0:65536 > gpurun_out/probe11.log 2>&1
cat gpurun_out/probe11.log
