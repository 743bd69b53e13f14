#!/bin/bash
# GPU session: parity tests + solver A/B on the contact configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() { # name, env assignments...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --config $CFG --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${CFG}_$name.json 2> gpurun_out/ab_${CFG}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_${CFG}_$name.json').read().strip().splitlines()[-1])
    print('$CFG $name: value %.4g  ms/step %.4g  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as ex:
    print('$CFG $name: parse failed', ex)
PY
}
for CFG in r2d2_maze from_the_readme basic_env ur_high_5; do
  run legacy DG_SOLVER=0
  run rs_auto DG_SOLVER=1
  run rs_global DG_SOLVER=1 DG_RS_ASHARED=0
done
CFG=r2d2_maze
run rs_a1024 DG_SOLVER=1 DG_RS_ASHARED=1024
run rs_a2048 DG_SOLVER=1 DG_RS_ASHARED=2048
