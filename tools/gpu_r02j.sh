#!/usr/bin/env bash
# profiles: launch list + ncu --set full of every stage at DTU, and the binning A/B (bins / radix / cub) as JSON
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 33 -c 33 --csv --log-file gpurun_out/r02j_launches_native_dtu.csv python tools/gpu_step.py native dtu 6 > gpurun_out/r02j_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 22 -c 11 -o gpurun_out/r02j_dtu_full python tools/gpu_step.py native dtu 3 > gpurun_out/r02j_ncu.log 2>&1
for cfg in dtu lego fern_pair; do
  for mode in bins radix cub; do
    B3GS_BINNING=$mode timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/r02j_binning_${cfg}_${mode}.json 2> gpurun_out/r02j_err.log
  done
done
python - <<'PY'
import json
res={}
for cfg in ("dtu","lego","fern_pair"):
    for mode in ("bins","radix","cub"):
        try:
            d=json.load(open("gpurun_out/r02j_binning_%s_%s.json"%(cfg,mode)))
            res["%s/%s"%(cfg,mode)]={"ms_per_step":d["ms_per_step"],"depth_sort_ms":d["kernels"]["depth_sort"]["ms"],"binning_ms":d["kernels"]["binning"]["ms"],"views_per_s":d["value"]}
        except Exception as ex: res["%s/%s"%(cfg,mode)]={"error":repr(ex)}
json.dump(res,open("gpurun_out/r02j_binning_ab.json","w"),indent=1); print(json.dumps(res,indent=1))
PY
ls -la gpurun_out/r02j*
