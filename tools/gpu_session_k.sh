#!/bin/bash
# scale sweep on one B200 (config 5 of BASELINE.json and the contact configs), final round-1 build
mkdir -p gpurun_out
python - <<'PY'
import json, subprocess, sys
out = {}
for cfg, sizes in (('ur_high_5_randomised', (1024, 2048, 4096, 8192, 16384, 32768, 65536)), ('r2d2_maze', (1024, 4096, 16384, 65536)),
                   ('basic_env', (1024, 4096, 16384, 65536)), ('ur_gripper', (1024, 4096, 16384))):
    for n in sizes:
        r = subprocess.run([sys.executable, 'bench.py', '--config', cfg, '--envs', str(n), '--steps', '20', '--warmup', '5', '--no-cpu-baseline'], capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            out['%s:%d' % (cfg, n)] = {'env_steps_per_s': d['value'], 'ms_per_step': d['ms_per_step'], 'kernel_ms': d['roofline']['kernel_ms'], 'team': d['config'].get('team'), 'grid_blocks': d['config'].get('grid_blocks')}
            print(cfg, n, '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'], flush=True)
        except Exception as e:
            print(cfg, n, 'failed', e, r.stderr[-300:], flush=True)
json.dump(out, open('gpurun_out/scale_sweep_1gpu_v2.json', 'w'), indent=1, sort_keys=True)
PY
