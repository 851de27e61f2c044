#!/bin/bash
# Launch list + one `--set full` capture of the longest launch of each heavy kernel, for one search.
# usage (on the GPU box): tools/gpu_profile.sh <tag> <ries args...>      outputs under gpurun_out/
set -u
tag=$1; shift
out=gpurun_out
mkdir -p $out
export RIES_NO_GRAPHS=1   # same kernels, launched directly, so that every launch is one ncu record
./ries_b200/bin/ries "$@" > $out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
  ./ries_b200/bin/ries "$@" > $out/${tag}_ncu1.log 2>&1
python tools/ncu_summary.py $out/${tag}_launches.csv | tee $out/${tag}_launch_summary.txt
python - "$out/${tag}_launches.csv" > $out/${tag}_pick.txt <<'PY'
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
best = {}
for i, row in enumerate(csv.DictReader(lines)):
    name = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
    if name not in best or v > best[name][1]:
        best[name] = (int(row["ID"]), v)
import os
for name in os.environ.get("KERNELS", "k_expand k_keys k_rs_onesweep k_classify_compact k_match_merge_new k_merge_old k_tiny_front").split():
    if name in best:
        print(name, best[name][0])
PY
cat $out/${tag}_pick.txt
# DRAM traffic of every launch of the dominant kernel (for bench.py's roofline.traffic)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:k_expand|k_tiny_front' --csv \
  --log-file $out/${tag}_k_expand_dram.csv ./ries_b200/bin/ries "$@" > $out/${tag}_ncu2.log 2>&1
python - "$out/${tag}_k_expand_dram.csv" <<'PY' | tee $out/${tag}_k_expand_traffic.json
import csv, json, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = 0.0; ids = set()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1)
    tot += v; ids.add(row["ID"])
print(json.dumps({"kernel": "k_expand + k_tiny_front (every K1 launch)", "launches": len(ids), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / max(1, len(ids))}))
PY
while read name idx; do
  timeout 600 ncu --set full --import-source on --clock-control none -s $idx -c 1 -f -o $out/${tag}_${name} \
    ./ries_b200/bin/ries "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ls -la $out/${tag}_${name}.ncu-rep
done < $out/${tag}_pick.txt
