#!/bin/bash
# multi-GPU bench lines (one process per GPU, no data-path collective): bash tools/gpu_session_h.sh <N>
N=$1
mkdir -p gpurun_out/h
for cfg in r2d2_maze ur_high_5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $cfg --no-cpu-baseline > gpurun_out/h/bench_${cfg}_${N}gpu.json 2> gpurun_out/h/bench_${cfg}_${N}gpu.err
  echo "$cfg x$N rc=$?"; tail -1 gpurun_out/h/bench_${cfg}_${N}gpu.json | cut -c1-260
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 5 --warmup 3 > gpurun_out/h/bench_reference_arm_${N}gpu.json 2> gpurun_out/h/bench_reference_arm_${N}gpu.err
echo "reference x$N rc=$?"; tail -1 gpurun_out/h/bench_reference_arm_${N}gpu.json | cut -c1-200
