#!/bin/bash
mkdir -p gpurun_out
run() { # cfg rs_min
  out=$(DG_SOLVER=1 DG_RS_ASHARED=0 DG_RS_MIN=$2 timeout 300 python bench.py --config $1 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$1" "$2" "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[3]); print('%s rs_min %s: value %.4g ms/step %.4g' % (sys.argv[1], sys.argv[2], d['value'], d['ms_per_step']))
except Exception as ex:
    print(sys.argv[1:3], 'failed', ex)
PY
}
for cfg in r2d2_maze from_the_readme basic_env; do
  for m in 0 10 13 19 25 37 1000; do run $cfg $m; done
done | tee gpurun_out/rs_min_sweep.log
