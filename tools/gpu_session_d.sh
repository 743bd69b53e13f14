#!/bin/bash
# team-size / block-size sweep of the row-space solver on the contact configs
mkdir -p gpurun_out
run() { # cfg team epb
  out=$(DG_SOLVER=1 DG_RS_ASHARED=0 DG_ENVS_PER_BLOCK=$3 timeout 300 python bench.py --config $1 --team $2 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$1" "$2" "$3" "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[4]); print('%s team %s epb %s: value %.4g ms/step %.4g' % (sys.argv[1], sys.argv[2], sys.argv[3], d['value'], d['ms_per_step']))
except Exception as ex:
    print(sys.argv[1:4], 'failed', ex)
PY
}
for cfg in r2d2_maze from_the_readme; do
  for team in 4 8 16 32; do for epb in 4 8 16; do run $cfg $team $epb; done; done
done | tee gpurun_out/team_sweep_rs.log
