#!/usr/bin/env bash
mkdir -p gpurun_out
for cfg in lego fern; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 12 --csv --log-file gpurun_out/r02r_launches_$cfg.csv python tools/gpu_step.py native $cfg 6 > /dev/null 2>&1
python - $cfg <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/r02r_launches_%s.csv'%sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
print(sys.argv[1], [(r[ki][:28], round(float(r[vi].replace(',',''))/1000,1)) for r in rows[1:]])
PY
done
