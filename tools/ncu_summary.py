"""Summarise an `ncu --set full` report of one rasterizer step into the tracked tables:
  profiles/<name>_summary.json  selected metrics per kernel launch
  profiles/traffic.json[workload]      DRAM bytes per launch of each stage (roofline.traffic in bench.py)
  profiles/issue_slots.json[workload]  warp instructions per launch of each stage (issue_slot_frac)
usage: python tools/ncu_summary.py gpurun_out/r02t_dtu_full.ncu-rep dtu profiles/r02t_dtu_full_summary.json
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size"]
STAGES = (("depth_sort", "depth_sort"), ("tile_bins_kernel", "binning"), ("bin_chunk_sums", "binning"),
          ("bin_tile_scan", "binning"), ("bin_apply", "binning"), ("composite_forward", "composite_forward"),
          ("composite_backward", "composite_backward"), ("preprocess_backward", "preprocess_backward"),
          ("preprocess_kernel", "preprocess"))


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value) * scale


def main(report, workload, out_path):
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units, launches = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    summary, traffic, slots, seen = [], {}, {}, set()
    for r in launches:
        name = r[col["Kernel Name"]]
        entry = {"Kernel Name": name, "units": {}}
        for k in KEEP:
            if k in col:
                entry[k] = r[col[k]]
                entry["units"][k] = units[col[k]]
        summary.append(entry)
        stage = next((s for key, s in STAGES if key in name), None)
        if stage is None or (name, stage) in seen:
            continue   # one launch of each kernel per stage (the report may span more than one step)
        seen.add((name, stage))
        dram = sum(to_bytes(r[col[k]], units[col[k]]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic[stage] = traffic.get(stage, 0) + int(round(dram))
        slots[stage] = slots.get(stage, 0) + int(float(r[col["smsp__inst_executed.sum"]]))
    json.dump(summary, open(out_path, "w"), indent=1)
    for fname, table in (("traffic.json", traffic), ("issue_slots.json", slots)):
        path = os.path.join(ROOT, "profiles", fname)
        doc = json.load(open(path))
        doc[workload] = table
        doc["_source_" + workload] = os.path.basename(out_path)
        json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps({"traffic": traffic, "issue_slots": slots}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
