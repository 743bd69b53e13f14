for c in r2d2_maze ur_high_5 from_the_readme drone_pilot ur_high_5_randomised basic_env; do python bench.py --config $c --steps 30 --warmup 5 > gpurun_out/bench1_$c.json 2> gpurun_out/bench1_$c.err; tail -c 300 gpurun_out/bench1_$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench1_ref.json 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
