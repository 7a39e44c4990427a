#!/bin/bash
# CSR (scaled-shape) path: parity tests + the configs[3] bench line, FP32 and tensor path
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sparse or csr or wide" 2>&1 | tail -5
python bench.py --workload fwd4096 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/fwd4096.json 2> gpurun_out/fwd4096.err
tail -3 gpurun_out/fwd4096.err
python - <<'PY'
import json
for l in open('gpurun_out/fwd4096.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'kernel_ms',d.get('roofline',{}).get('kernel_ms'),'frac',d.get('roofline',{}).get('path'))
PY
