#!/bin/bash
# BASELINE.json config 5 sweep (ur_high_5 with per-environment randomised masses / friction / damping / poses) and the headline
# configs at 1 / 2 / 4 / 8 GPUs of one box:   gpurun --gpus 8 -- 'bash tools/scale_sweep.sh'
# One process per GPU (torchrun), environments sharded, no per-step collective; every line is bench.py's own JSON line.
OUT=gpurun_out/${TAG:-r2_scale}
mkdir -p $OUT
run() {  # gpus config envs
  local g=$1 cfg=$2 n=$3 f=$OUT/${2}_${3}envs_${1}gpu.json
  if [ $g -eq 1 ]; then python bench.py --gpus 1 --config $cfg --envs $n --steps 20 --warmup 5 --preroll 100 --no-cpu-baseline --no-per-config > $f 2> $f.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $g --config $cfg --envs $n --steps 20 --warmup 5 --preroll 100 --no-cpu-baseline --no-per-config > $f 2> $f.err; fi
  python - <<PY
import json
try:
    d = json.loads(open('$f').read().strip().splitlines()[-1])
    print('%-22s envs/GPU %6d  GPUs %d  value %.4g env-steps/s  ms/step %.4g  e2e %.4g' % ('$cfg', $n, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value']))
except Exception as e:
    print('$cfg $n $g FAILED', e)
PY
}
for g in ${GPUS:-1 2 4 8}; do
  for n in ${SIZES:-1024 4096 16384 65536}; do run $g ur_high_5_randomised $n; done
  run $g r2d2_maze 4096
  run $g ur_high_5 8192
done
