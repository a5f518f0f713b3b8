#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   -- bench.py (device-resident arm only) once per value of an environment knob
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['roofline']['phase_ms']
print('$var=$v', 'ms/step %.3f' % d['ms_per_step'], ' '.join('%s %.3f' % (k.split('(')[0], x) for k, x in p.items()), 'e2e %.2f' % d['e2e']['ms_per_step'])"
done
