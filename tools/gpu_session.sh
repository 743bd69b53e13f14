#!/bin/bash
# One documented GPU session script (replaces the round-1 gpu_session_[a-k].sh scratch files).  Runs on the B200 box through
#   gpurun --timeout <s> -- 'bash tools/gpu_session.sh <stage> [...]'
# and writes everything under gpurun_out/<tag>/ (merged back into the build container).  Stages:
#   tests            pytest -m gpu (log only)
#   qd               tools/qd_probe.py, SFU and libm sincos
#   bench <cfg...>   bench.py --config <cfg> (no CPU baseline), default solver and DG_RS_MIN A/B when RS_AB=1
#   phase <cfg...>   tools/phase_probe.py
#   ncu_step <cfg>   launch list + one --set full capture of dg_step_kernel
#   ncu_render       launch list + one --set full capture of dg_render_kernel (from_the_readme)
TAG=${TAG:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
stage=$1; shift
case $stage in
  tests)
    python -m pytest tests -m gpu -q -x --timeout 1200 "$@" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log ;;
  qd)
    python tools/qd_probe.py --envs 64 --steps 10 > $OUT/qd_probe_sfu.json 2> $OUT/qd_probe_sfu.err
    DG_PRECISE=1 python tools/qd_probe.py --envs 64 --steps 10 > $OUT/qd_probe_libm.json 2> $OUT/qd_probe_libm.err
    grep -A10 worst_over_steps $OUT/qd_probe_sfu.json | head -24; grep -A10 worst_over_steps $OUT/qd_probe_libm.json | head -24 ;;
  bench)
    for cfg in "$@"; do
      python bench.py --config $cfg --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline > $OUT/bench_${cfg}${SUFFIX}.json 2> $OUT/bench_${cfg}${SUFFIX}.err
      python - <<PY
import json
try:
    d = json.loads(open('$OUT/bench_${cfg}${SUFFIX}.json').read().strip().splitlines()[-1])
    print('$cfg$SUFFIX', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'kernel_ms %.4g' % d['roofline']['kernel_ms'], 'ms/step %.4g' % d['ms_per_step'], d['config'].get('team'), d['config'].get('block_threads'))
except Exception as e:
    print('$cfg$SUFFIX', 'FAILED', e); print(open('$OUT/bench_${cfg}${SUFFIX}.err').read()[-2000:])
PY
    done ;;
  phase)
    for cfg in "$@"; do python tools/phase_probe.py $cfg ${ENVS:-} > $OUT/phase_${cfg}${SUFFIX}.log 2>&1; head -14 $OUT/phase_${cfg}${SUFFIX}.log; done ;;
  ncu_step)
    cfg=$1
    # one --set full capture of a stage launch of the step kernel after pre-roll (3 stage launches per step under the split schedule;
    # SKIP picks which: 901 = [post + pre] of step 300, the robots are on the ground by then)
    ncu --set full --clock-control none --import-source on -k regex:dg_step_kernel -s ${SKIP:-901} -c 1 -o $OUT/ncu_full_step_${cfg} -f python tools/profile_cmd.py $cfg 0 ${ENVS:-4096} 320 > $OUT/ncu_full_step_${cfg}.log 2>&1
    ncu -i $OUT/ncu_full_step_${cfg}.ncu-rep --page details > $OUT/ncu_details_step_${cfg}.txt 2>&1
    grep -E "Duration|Registers Per|Achieved Occ|Executed Ipc|L1/TEX Hit|Active Warps Per|Avg. Active Threads" $OUT/ncu_details_step_${cfg}.txt
    ;;  # by source function: python tools/ncu_by_func.py <rep> diy_gym_b200/libdiygym_b200.so dg_step_kernelILi8 (run where the .so was built)
  ncu_solve)
    cfg=$1
    # launch list of 10 steps after 300 steps of pre-roll (7 kernel launches per step under the split schedule: 3 stage + 4 sweep)
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${SKIP:-2100} -c 70 --csv --log-file $OUT/ncu_launches_${cfg}.csv python tools/profile_cmd.py $cfg 0 ${ENVS:-4096} 320 > $OUT/ncu_launches_${cfg}.log 2>&1
    python tools/ncu_launch_summary.py $OUT/ncu_launches_${cfg}.csv 10 > $OUT/ncu_launches_${cfg}_summary.txt 2>&1; cat $OUT/ncu_launches_${cfg}_summary.txt
    # the sweep kernel is launched twice per sub-step: rows-per-lane K = 2 first (auxiliary stream), then K = 1
    ncu --set full --clock-control none --import-source on -k regex:dg_solve_kernel -s 1200 -c 2 -o $OUT/ncu_full_solve_${cfg} -f python tools/profile_cmd.py $cfg 0 ${ENVS:-4096} 320 > $OUT/ncu_full_solve_${cfg}.log 2>&1
    ncu -i $OUT/ncu_full_solve_${cfg}.ncu-rep --page details > $OUT/ncu_details_solve_${cfg}.txt 2>&1 ;;
  ncu_render)
    ncu --set full --clock-control none --import-source on -k regex:dg_render_kernel -s 2 -c 1 -o $OUT/ncu_full_render -f python tools/render_probe.py from_the_readme ${ENVS:-4096} > $OUT/ncu_full_render.log 2>&1
    ncu -i $OUT/ncu_full_render.ncu-rep --page details > $OUT/ncu_details_render.txt 2>&1
    ncu -i $OUT/ncu_full_render.ncu-rep --page raw --csv > $OUT/ncu_raw_render.csv 2>&1 ;;
  *) echo "unknown stage $stage"; exit 2 ;;
esac
