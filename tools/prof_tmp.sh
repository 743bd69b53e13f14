python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu11.log; cat gpurun_out/pytest_gpu11.log
timeout 600 python tools/gpu_probe3.py ur_high_5:0:0:3 r2d2_maze:0:0:3 r2d2_maze:8:8:3 from_the_readme:0:0:3 from_the_readme:8:8:3 drone_pilot:0:0:3 drone_pilot:4:8:3 basic_env:0:0:3 ur_high_5:0:0:3:65536 ur_high_5:0:0:3:1024 r2d2_maze:0:0:3:16384 2>&1 | tee gpurun_out/probe27.log
