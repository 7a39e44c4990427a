#!/bin/bash
# quick iteration loop: smoke parity + stage timings
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $BENCH_ARGS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'kernel_ms', {k: round(v,3) for k,v in r['kernel_ms'].items()}, 'inproj_frac', round(r['frac'],3), 'path_frac', round(r['path']['frac_fp32'],3), 'peak', round(r['peak'],1))
    else: print(l)
"
