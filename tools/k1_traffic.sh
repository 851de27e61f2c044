export RIES_NO_GRAPHS=1
out=gpurun_out; tag=r1l6
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:k_expand|k_tiny_front' --csv --log-file $out/${tag}_k_expand_dram.csv ./ries_b200/bin/ries 2.5063141592653589 -l6 -F3 --no-slow-messages > $out/${tag}_ncu2.log 2>&1
python - "$out/${tag}_k_expand_dram.csv" <<'PY' | tee $out/${tag}_k_expand_traffic.json
import csv, json, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = 0.0; ids = set()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1)
    tot += v; ids.add(row["ID"])
print(json.dumps({"kernel": "k_expand + k_tiny_front (every K1 launch)", "launches": len(ids), "dram_bytes_total": tot, "dram_bytes_per_launch": tot / max(1, len(ids))}))
PY
