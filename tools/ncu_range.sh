#!/bin/bash
# usage: tools/ncu_range.sh [envs] [launches] -> gpurun_out/r2_range_<envs>.csv + profiles/r2_step_tma_range.json
n=${1:-1048576}; l=${2:-16}
mkdir -p gpurun_out
ncu --replay-mode app-range --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum \
    --csv --log-file gpurun_out/r2_range_$n.csv python tools/ncu_range.py $n $l > gpurun_out/r2_range_$n.log 2>&1
python - "$n" "$l" <<'PY'
import csv, json, sys
n, l = int(sys.argv[1]), int(sys.argv[2])
vals = {}
with open(f"gpurun_out/r2_range_{n}.csv") as f:
    rows = [r for r in csv.reader(f) if len(r) > 5]
hdr = rows[0]
mi, vi, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1}
for r in rows[1:]:
    vals[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
out = {"envs_per_launch": n, "launches": l, "dram_bytes_read": vals["dram__bytes_read.sum"],
       "dram_bytes_write": vals["dram__bytes_write.sum"], "range_ns": vals.get("gpu__time_duration.sum"),
       "algorithmic_bytes_per_launch": 117 * n,
       "how": f"ncu --replay-mode app-range over {l} consecutive k_step_tma launches of the bench ring "
              f"(tools/ncu_range.sh); bytes = (dram__bytes_read.sum + dram__bytes_write.sum) / {l}"}
name = "profiles/r2_step_tma_range.json" if n == 1048576 else f"profiles/r2_step_tma_range_{n}.json"
json.dump(out, open("gpurun_out/" + name.split("/")[1], "w"), indent=1)
print(json.dumps(out))
PY
