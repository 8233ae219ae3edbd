#!/usr/bin/env python
"""profiles/r02_traffic.json from an ncu report: per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of
the step's kernels, tagged with the configuration and a hash of the kernel sources the capture was taken from --
bench.py reports `roofline.traffic` only while that hash matches the tree.

    ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/make_traffic.py /tmp/raw.csv 8192 128 8 "<command the capture ran>"
"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
n, T, r = (int(x) for x in sys.argv[2:5])
out = {"n": n, "T": T, "r": r, "source_hash": bench.kernel_source_hash(), "command": sys.argv[5] if len(sys.argv) > 5 else None,
       "metric": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full --clock-control none)"}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
for rw in rows[2:]:
    name = rw[ix["Kernel Name"]].split("<")[0].replace("void ", "").strip()
    rd = float(rw[ix["dram__bytes_read.sum"]]) * scale[units[ix["dram__bytes_read.sum"]]]
    wr = float(rw[ix["dram__bytes_write.sum"]]) * scale[units[ix["dram__bytes_write.sum"]]]
    out[name] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                 "duration_ms": float(rw[ix["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0),
                 "kernel": rw[ix["Kernel Name"]]}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
