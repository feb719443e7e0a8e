#!/bin/bash
# quick GPU check used during kernel work: parity tests selected by $1 (pytest -k), then one bench line per workload in $2..
K="$1"; shift
python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -4
for w in "$@"; do
  python bench.py --workload $w --no-cpu-baseline --no-e2e --also none --sustain 0 --steps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w', 'ms/step %.4f' % d['ms_per_step'], 'value %.4g' % d['value'], 'frac %.4f' % r['frac'], {k: round(v,4) for k,v in r['kernels_ms_per_step'].items()}, d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
