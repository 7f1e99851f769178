"""Write profiles/r2_traffic.json from an ncu --set full capture of the dominant kernel: DRAM bytes read + written by that
launch, stamped with the hash of the kernel sources (bench.py reports `roofline.traffic` only while the hash matches)."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_sha only; nothing is run)

rep, what = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]


def metric(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return float(vals[i]) * scale


rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
out = {"traffic": int(rd + wr), "unit": "bytes", "dram_read": int(rd), "dram_write": int(wr), "what": what,
       "kernel": vals[hdr.index("Kernel Name")], "duration_us_under_ncu": float(vals[hdr.index("gpu__time_duration.sum")]),
       "source_sha": bench.source_sha(), "capture": rep}
for path in (os.path.join(ROOT, "profiles", "r2_traffic.json"), os.path.join(ROOT, "gpurun_out", "r2_traffic.json")):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
