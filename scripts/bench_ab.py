"""A/B of environment switches through bench.py:  python scripts/bench_ab.py A=1 A=0 A=1,B=16 [-- extra bench args]
(each argument: one run with that comma-separated set of variables).  Prints value / ms_per_step / e2e and the stage medians of each run (one bench.py process per setting)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
extra = []
if '--' in args:
    extra = args[args.index('--') + 1:]
    args = args[:args.index('--')]
for setting in args:
    name, val = setting, ''
    env = dict(os.environ, **dict(kv.split('=', 1) for kv in setting.split(',') if kv))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '60', '--warmup', '5', *extra],
                       env=env, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith('{')]
    if not lines:
        print(f'{name}{val}: bench failed rc={r.returncode}\n{r.stderr[-2000:]}')
        continue
    j = json.loads(lines[-1])
    stages = {k: v for k, v in j.items() if 'stage' in k}
    print(f"{name}{val}: value {j['value']:.0f} {j['unit']}, {j['ms_per_step'] * 1e3:.2f} us/step, e2e {j['e2e']['value']:.0f}; {stages}")
