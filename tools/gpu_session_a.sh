#!/bin/bash
# GPU session: parity tests, then A/B of the contact solver (DG_SOLVER=0 dv-space per body, 1 row-space team) per config
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for cfg in r2d2_maze ur_high_5 basic_env from_the_readme; do
  for s in 0 1; do
    DG_SOLVER=$s timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_${cfg}_solver$s.json 2> gpurun_out/ab_${cfg}_solver$s.err
    echo "$cfg solver=$s rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/ab_${cfg}_solver$s.json').read().strip().splitlines()[-1])
    print('  value %.4g  ms/step %.4g  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as ex:
    print('  parse failed', ex)
PY
  done
done
