"""profiles/r02_ncu_traffic.json from ncu --set full reports: DRAM bytes per launch of the kernels bench.py may report as dominant,
stamped with the sha256 of the CUDA sources the capture was taken from (bench.py quotes `roofline.traffic` only when the stamp
matches the sources it is running).   python profiles/make_traffic_json.py name=path.ncu-rep ...   (run in the build container)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_hash

out = {"source_hash": source_hash(), "kernels": {}}
for arg in sys.argv[1:]:
    name, path = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    def val(r, key):
        i = hdr.index(key)
        v = float(r[i].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3}.get(units[i], 1)
    rs = rows[2:]
    out["kernels"][name] = {
        "dram_bytes_read": sum(val(r, "dram__bytes_read.sum") for r in rs) / len(rs),
        "dram_bytes_write": sum(val(r, "dram__bytes_write.sum") for r in rs) / len(rs),
        "duration_us": sum(val(r, "gpu__time_duration.sum") for r in rs) / len(rs),
        "grid": rs[0][hdr.index("Grid Size")], "launches": len(rs),
        "note": f"ncu --set full, {os.path.basename(path)}"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
