#!/bin/bash
mkdir -p gpurun_out
{
echo "Per-phase cycle breakdown of one step-kernel launch (dg_debug_phase_cycles, tools/phase_probe.py): thread 0 of every block sums the"
echo "cycles of each phase, barrier included; ms at 1.965 GHz; launch 60..63 of a rollout with random actions.  The line named"
echo "phase_rs_build is the solver phase of the launch (row-space build, or the per-body sweeps pgs_body_full / pgs_unit)."
for cfg in r2d2_maze ur_high_5 from_the_readme basic_env ur_gripper ur_admittance; do echo; python tools/phase_probe.py $cfg 2>&1 | tail -22; done
echo; echo "=== r2d2_maze with every contact environment in row space (DG_RS_MIN=0)"; DG_RS_MIN=0 python tools/phase_probe.py r2d2_maze 2>&1 | tail -22
} > gpurun_out/phase_probe.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline | tail -1 | cut -c1-200
