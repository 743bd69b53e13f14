python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/plain_bench2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1
rm -f profiles/ncu_traffic.json
for c in r2d2_maze:4096 ur_high_5:8192 from_the_readme:4096 drone_pilot:4096 ur_high_5_randomised:8192 basic_env:4096; do n=${c%%:*}; e=${c##*:}; ncu --set full --clock-control none -k regex:dg_step_kernel -s 20 -c 1 -f -o /tmp/prof_$n python tools/profile_cmd.py $n 0 $e 21 > gpurun_out/ncu_p_$n.log 2>&1; tail -n 1 gpurun_out/ncu_p_$n.log; python tools/ncu_traffic.py $c=/tmp/prof_$n.ncu-rep; done
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
ncu -i /tmp/prof_r2d2_maze.ncu-rep --page details > gpurun_out/ncu_details_r2d2_maze.txt 2>&1
ncu -i /tmp/prof_ur_high_5.ncu-rep --page details > gpurun_out/ncu_details_ur_high_5.txt 2>&1
du -sh gpurun_out
