#!/bin/bash
# end-to-end pipeline variants of bench.py: slices of the batch x device buffers per slice
for cb in "1 1" "1 2" "2 1" "2 2"; do
  set -- $cb
  python bench.py --e2e-chunks $1 --e2e-buffers $2 --lockstep-iters 0 --eval-points 0 --no-descent --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunks $1 buffers $2:', round(d['value']), 'device builds/s;', round(d['e2e']['value']), 'e2e builds/s;', round(d['e2e']['ms_per_step'],3), 'ms; ok', d['e2e']['builds_ok'])"
done
