#!/bin/bash
# DRAM traffic of one scan-kernel launch over the 1 GiB bench workload, tied to the library build:
# writes profiles/scan_kernel_traffic.json (bench.py reports it only for the same scan-kernel sources + flags, or the same .so).
mkdir -p gpurun_out
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:scan_kernel -s 2 -c 1 --csv --log-file gpurun_out/scan_traffic.csv python scripts/ncu_target.py > gpurun_out/scan_traffic.log 2>&1
python - <<'PY'
import csv, hashlib, json, sys
sys.path.insert(0, ".")
import bench
from pathlib import Path
rows = [r for r in csv.reader(open("gpurun_out/scan_traffic.csv")) if len(r) > 10 and r[0].isdigit()]
vals = {r[-3]: (float(r[-1].replace(",", "")), r[-2]) for r in rows}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = vals["dram__bytes_read.sum"][0] * scale[vals["dram__bytes_read.sum"][1]]
wr = vals["dram__bytes_write.sum"][0] * scale[vals["dram__bytes_write.sum"][1]]
so = hashlib.sha256(Path("dump1090_b200/libmodes_b200.so").read_bytes()).hexdigest()[:16]
doc = {"kernel": "scan_kernel", "workload": "modes1.bin tiled to 1 GiB (one launch)", "dram_bytes_read": rd, "dram_bytes_write": wr,
       "dram_bytes_per_launch": rd + wr, "algorithmic_bytes_per_launch": 2 ** 30, "ratio": round((rd + wr) / 2 ** 30, 4),
       "so_sha256_16": so, "scan_source_sha256_16": bench.scan_source_digest(), "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one launch, scripts/ncu_traffic.sh"}
Path("gpurun_out/scan_kernel_traffic.json").write_text(json.dumps(doc, indent=1))
print(doc)
PY
