#!/bin/bash
# One `ncu --set full` capture of ONE launch of the persistent ADMM kernel on the bench workload (run on a GPU box, one
# GPU): writes gpurun_out/prof_admm_$TAG.ncu-rep, its raw / details pages, and gpurun_out/traffic_$TAG.json =
# dram__bytes_read.sum + dram__bytes_write.sum of that launch.  Copy the json to profiles/r2_traffic.json (bench.py
# reads `roofline.traffic` from there; a number taken under the profiler is never a bench value).
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
B200QP_AUTOTUNE=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:admm_kernel -c 1 -f -o gpurun_out/prof_admm_$TAG \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-spmv --no-batch --no-structured > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_admm_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_admm_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_admm_$TAG.ncu-rep --page details > gpurun_out/prof_admm_${TAG}_details.txt 2>/dev/null
ncu -i gpurun_out/prof_admm_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_admm_${TAG}_source.csv 2>/dev/null
python - <<PY
import csv, json
rows = list(csv.reader(open("gpurun_out/prof_admm_${TAG}_raw.csv")))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
def get(name):
    v = float(vals[col[name]].replace(",", "")); u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
out = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
       "kernel": vals[col["Kernel Name"]], "duration_under_ncu": vals[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]],
       "source": "ncu --set full --clock-control none, one launch of the persistent ADMM kernel on the bench workload "
                 "(scripts/ncu_traffic.sh; raw page: profiles/r2_admm_kernel_ncu_raw.csv)"}
json.dump(out, open("gpurun_out/traffic_$TAG.json", "w"), indent=1)
print(out)
PY
