#!/bin/bash
# A/B the occupancy variants of the composite kernels (one GPU call)
for f in 3 4 5 6; do for b in 3 4 5 6; do
  if [ $f -ne 4 ] && [ $b -ne 4 ]; then continue; fi
  B3GS_FWD_OCC=$f B3GS_BWD_OCC=$b timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fwd_occ=$f bwd_occ=$b', d['value'], d['kernels']['composite_forward']['ms'], d['kernels']['composite_backward']['ms'])"
done; done
